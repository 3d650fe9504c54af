// Opcode table of the elimination plan. Keep in sync with graphax_b200/lowering.py:OPCODES
// (tests/test_plan.py compares the two).
#ifndef GX_OPS_H
#define GX_OPS_H

#define GX_OPCODE_LIST(X) \
  X(INPUT, input) X(CONST, const) \
  X(ADD, add) X(SUB, sub) X(MUL, mul) X(DIV, div) X(FMA, fma) X(NEG, neg) X(ABS, abs) X(SQRT, sqrt) \
  X(SIN, sin) X(COS, cos) X(TAN, tan) X(ATAN2, atan2) X(EXP, exp) X(LOG, log) X(LOG1P, log1p) \
  X(TANH, tanh) X(SINH, sinh) X(COSH, cosh) X(ASIN, asin) X(ACOS, acos) X(ATAN, atan) \
  X(ASINH, asinh) X(ACOSH, acosh) X(ATANH, atanh) X(ERF, erf) X(POW, pow) X(LOGISTIC, logistic) \
  X(STORE_JAC, store_jac) X(STORE_PRIMAL, store_primal) \
  X(MAX, max) X(MIN, min) X(GT, gt) X(LT, lt) X(EQ, eq) \
  X(SELECT, select)

enum gx_opcode {
#define GX_ENUM(NAME, name) GX_OP_##NAME,
  GX_OPCODE_LIST(GX_ENUM)
#undef GX_ENUM
  GX_OP_COUNT
};

// operand = kind << 14 | index
#define GX_K_SLOT 0u
#define GX_K_CONST 1u
#define GX_K_INPUT 2u
#define GX_OPERAND_KIND(x) ((unsigned)(x) >> 14)
#define GX_OPERAND_INDEX(x) ((unsigned)(x) & 0x3fffu)

#define GX_BLOB_MAGIC 0x4C505847u
#define GX_BLOB_VERSION 1u
#define GX_FLAG_HAS_PRIMAL 1u

struct gx_blob_header {
  unsigned magic, version, total_bytes, n_in_arrays, n_out_arrays, n_in_scalars, n_cols, n_rows,
      n_ops, n_slots, n_consts, flags, reserved[4];
};

struct gx_op {
  unsigned short op, dst, a, b, c, pad;
};

// select: dst = a != 0 ? b : c   (lax.select_n(which, case0, case1) with b = case1, c = case0)
#define GX_OP_IS_TERNARY(op) ((op) == GX_OP_FMA || (op) == GX_OP_SELECT)

#endif
