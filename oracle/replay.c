/*
 * ORACLE - test infrastructure, not product code.
 *
 * CPU replay of a graphax_b200 elimination-plan blob (layout: graphax_b200/plan.py).  It executes
 * exactly the flat program that stands for the reference's run-time work - elemental partials
 * (/root/reference/src/graphax/primitives.py:150-205), the products and fill-in additions of
 * _eliminate_vertex (core.py:155-272) and the densification (core.py:555-562) - one IEEE-754
 * binary32 operation per plan op, `fma` through fmaf().  The CUDA kernels must agree with it bit
 * for bit on every plan whose ops are correctly rounded (add, sub, mul, div, fma, neg, abs, sqrt);
 * libm-dependent ops (sin, cos, atan2, ...) agree to a few ulp.
 *
 * Written independently of csrc/gx_abi.cu; it shares only the blob specification.  Samples are
 * processed in blocks of LANES so that the compiler can vectorise each op across samples (this is
 * also the "fused CPU" baseline bench.py reports); threads split the batch statically.
 *
 * Build: see oracle/Makefile (gcc -O3 -mfma -ffp-contract=off; no -ffast-math).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LANES 16

enum {
  OP_INPUT, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_FMA, OP_NEG, OP_ABS, OP_SQRT, OP_SIN, OP_COS, OP_TAN,
  OP_ATAN2, OP_EXP, OP_LOG, OP_LOG1P, OP_TANH, OP_SINH, OP_COSH, OP_ASIN, OP_ACOS, OP_ATAN, OP_ASINH, OP_ACOSH,
  OP_ATANH, OP_ERF, OP_POW, OP_LOGISTIC, OP_STORE_JAC, OP_STORE_PRIMAL, OP_MAX, OP_MIN, OP_GT, OP_LT, OP_EQ, OP_SELECT, OP_COUNT
};

typedef struct {
  uint32_t magic, version, total_bytes, n_in_arrays, n_out_arrays, n_in_scalars, n_cols, n_rows, n_ops, n_slots,
      n_consts, flags, reserved[4];
} header_t;

typedef struct { uint16_t op, dst, a, b, c, pad; } op_t;

typedef struct {
  const header_t* h;
  const uint32_t* in_sizes;
  const float* consts;
  const op_t* ops;
  const float* const* in;
  float* jac;
  float* primal;
  int64_t begin, end;
  int in_arr[16384];   /* input scalar -> array */
  int in_comp[16384];  /* input scalar -> component */
} job_t;

static const char* opcode_names[OP_COUNT] = {
  "input", "const", "add", "sub", "mul", "div", "fma", "neg", "abs", "sqrt", "sin", "cos", "tan", "atan2", "exp",
  "log", "log1p", "tanh", "sinh", "cosh", "asin", "acos", "atan", "asinh", "acosh", "atanh", "erf", "pow",
  "logistic", "store_jac", "store_primal", "max", "min", "gt", "lt", "eq", "select"};

int gxo_opcode_count(void) { return OP_COUNT; }
const char* gxo_opcode_name(int i) { return i >= 0 && i < OP_COUNT ? opcode_names[i] : ""; }

/* operand -> pointer to LANES floats (slot row, broadcast constant row, or gathered input row) */
static inline const float* operand(const job_t* j, uint16_t o, float* slots, float* cbuf, const float* inbuf) {
  const unsigned kind = o >> 14, idx = o & 0x3fff;
  if (kind == 0) return slots + (size_t)idx * LANES;
  if (kind == 1) {
    for (int l = 0; l < LANES; ++l) cbuf[l] = j->consts[idx];
    return cbuf;
  }
  return inbuf + (size_t)idx * LANES;
}

#define UNARY(expr)                                  \
  for (int l = 0; l < LANES; ++l) { const float x = a[l]; d[l] = (expr); } break
#define BINARY(expr)                                 \
  for (int l = 0; l < LANES; ++l) { const float x = a[l], y = b[l]; d[l] = (expr); } break

static void run_block(const job_t* j, int64_t s0, int n, float* slots, float* inbuf) {
  const header_t* h = j->h;
  const unsigned tile = h->n_rows * h->n_cols;
  float ca[LANES], cb[LANES], cc[LANES];
  for (unsigned k = 0; k < h->n_in_scalars; ++k) {
    const int arr = j->in_arr[k], comp = j->in_comp[k];
    const unsigned sz = j->in_sizes[arr];
    for (int l = 0; l < LANES; ++l) inbuf[k * LANES + l] = l < n ? j->in[arr][(s0 + l) * sz + comp] : 1.0f;
  }
  for (unsigned i = 0; i < h->n_ops; ++i) {
    const op_t* o = &j->ops[i];
    const float* a = operand(j, o->a, slots, ca, inbuf);
    if (o->op == OP_STORE_JAC) {
      for (int l = 0; l < n; ++l) j->jac[(s0 + l) * tile + o->dst] = a[l];
      continue;
    }
    if (o->op == OP_STORE_PRIMAL) {
      if (j->primal) for (int l = 0; l < n; ++l) j->primal[(s0 + l) * h->n_rows + o->dst] = a[l];
      continue;
    }
    const float* b = a;
    const float* c = a;
    if (o->op == OP_ADD || o->op == OP_SUB || o->op == OP_MUL || o->op == OP_DIV || o->op == OP_FMA ||
        o->op == OP_ATAN2 || o->op == OP_POW || o->op == OP_MAX || o->op == OP_MIN || o->op == OP_GT ||
        o->op == OP_LT || o->op == OP_EQ || o->op == OP_SELECT)
      b = operand(j, o->b, slots, cb, inbuf);
    if (o->op == OP_FMA || o->op == OP_SELECT) c = operand(j, o->c, slots, cc, inbuf);
    float t[LANES];
    float* d = t; /* operands may alias the destination slot: compute into a temporary row */
    switch (o->op) {
      case OP_ADD: BINARY(x + y);
      case OP_SUB: BINARY(x - y);
      case OP_MUL: BINARY(x * y);
      case OP_DIV: BINARY(x / y);
      case OP_FMA: for (int l = 0; l < LANES; ++l) d[l] = fmaf(a[l], b[l], c[l]); break;
      case OP_NEG: UNARY(-x);
      case OP_ABS: UNARY(fabsf(x));
      case OP_SQRT: UNARY(sqrtf(x));
      case OP_SIN: UNARY(sinf(x));
      case OP_COS: UNARY(cosf(x));
      case OP_TAN: UNARY(tanf(x));
      case OP_ATAN2: BINARY(atan2f(x, y));
      case OP_EXP: UNARY(expf(x));
      case OP_LOG: UNARY(logf(x));
      case OP_LOG1P: UNARY(log1pf(x));
      case OP_TANH: UNARY(tanhf(x));
      case OP_SINH: UNARY(sinhf(x));
      case OP_COSH: UNARY(coshf(x));
      case OP_ASIN: UNARY(asinf(x));
      case OP_ACOS: UNARY(acosf(x));
      case OP_ATAN: UNARY(atanf(x));
      case OP_ASINH: UNARY(asinhf(x));
      case OP_ACOSH: UNARY(acoshf(x));
      case OP_ATANH: UNARY(atanhf(x));
      case OP_ERF: UNARY(erff(x));
      case OP_POW: BINARY(powf(x, y));
      case OP_LOGISTIC: UNARY(1.0f / (1.0f + expf(-x)));
      case OP_MAX: BINARY(fmaxf(x, y));
      case OP_MIN: BINARY(fminf(x, y));
      case OP_GT: BINARY(x > y ? 1.0f : 0.0f);
      case OP_LT: BINARY(x < y ? 1.0f : 0.0f);
      case OP_EQ: BINARY(x == y ? 1.0f : 0.0f);
      case OP_SELECT: for (int l = 0; l < LANES; ++l) d[l] = a[l] != 0.0f ? b[l] : c[l]; break;
      default: for (int l = 0; l < LANES; ++l) d[l] = 0.0f; break;
    }
    memcpy(slots + (size_t)o->dst * LANES, t, sizeof t);
  }
}

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  float* slots = (float*)aligned_alloc(64, sizeof(float) * LANES * (j->h->n_slots + 1));
  float* inbuf = (float*)aligned_alloc(64, sizeof(float) * LANES * (j->h->n_in_scalars + 1));
  for (int64_t s = j->begin; s < j->end; s += LANES) {
    const int n = (int)(j->end - s < LANES ? j->end - s : LANES);
    run_block(j, s, n, slots, inbuf);
  }
  free(slots);
  free(inbuf);
  return NULL;
}

/* returns 0 on success, <0 on a malformed blob */
int gxo_replay(const void* blob, size_t nbytes, int64_t batch, const float* const* in_ptrs, float* jac_out,
               float* primal_out, int n_threads) {
  if (!blob || nbytes < sizeof(header_t)) return -1;
  const header_t* h = (const header_t*)blob;
  if (h->magic != 0x4C505847u || h->version != 1u || h->total_bytes != nbytes) return -1;
  if (h->n_in_arrays > 32 || h->n_in_scalars > 0x3fff) return -1;
  const unsigned char* cur = (const unsigned char*)blob + sizeof(header_t);
  const uint32_t* in_sizes = (const uint32_t*)cur;
  cur += 4 * (size_t)(2 * h->n_in_arrays + h->n_out_arrays);
  const float* consts = (const float*)cur;
  cur += 4 * (size_t)h->n_consts;
  const op_t* ops = (const op_t*)cur;
  if ((size_t)(cur - (const unsigned char*)blob) + sizeof(op_t) * (size_t)h->n_ops != nbytes) return -1;
  for (unsigned i = 0; i < h->n_ops; ++i)
    if (ops[i].op >= OP_COUNT || ops[i].op < OP_ADD) return -2;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  if ((int64_t)n_threads * LANES > batch) n_threads = (int)((batch + LANES - 1) / LANES);
  if (n_threads < 1) n_threads = 1;

  job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof(job_t));
  pthread_t* tids = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  const int64_t blocks = (batch + LANES - 1) / LANES;
  for (int t = 0; t < n_threads; ++t) {
    job_t* j = &jobs[t];
    j->h = h; j->in_sizes = in_sizes; j->consts = consts; j->ops = ops;
    j->in = in_ptrs; j->jac = jac_out; j->primal = (h->flags & 1u) ? primal_out : NULL;
    j->begin = (blocks * t / n_threads) * LANES;
    j->end = (blocks * (t + 1) / n_threads) * LANES;
    if (j->end > batch) j->end = batch;
    unsigned k = 0;
    for (unsigned a = 0; a < h->n_in_arrays; ++a)
      for (unsigned c = 0; c < in_sizes[a]; ++c, ++k) { j->in_arr[k] = (int)a; j->in_comp[k] = (int)c; }
  }
  for (int t = 1; t < n_threads; ++t) pthread_create(&tids[t], NULL, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < n_threads; ++t) pthread_join(tids[t], NULL);
  free(jobs);
  free(tids);
  return 0;
}
