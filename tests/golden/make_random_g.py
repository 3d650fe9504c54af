"""Extract the op table of the reference's random scalar function ``g``.

``g`` (15 scalar inputs, 20 outputs, 105 equations) is the only function for
which the reference records literal known answers: the ``count_ops`` totals
``# fwd: 632, rev: 566, mM: 451`` (``src/graphax/examples/randoms.py:91``).
This script (run once, in the build container, where /root/reference exists)
turns ``randoms.py:92-200`` into a data table ``random_g.json`` so the tests can
rebuild the same computational graph without the reference being present.
"""
import json
import re
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/graphax/examples/randoms.py")
lines = src.read_text().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith("def g("))
args = re.findall(r"v\d+", lines[start])
ops, outputs = [], None
for l in lines[start + 1:]:
    l = l.strip()
    if l.startswith("return"):
        outputs = re.findall(r"v\d+", l)
        break
    m = re.match(r"(v\d+) = jnp\.(\w+)\((.*)\)$", l)
    assert m, l
    ops.append([m.group(1), m.group(2), [a.strip() for a in m.group(3).split(",")]])
out = {"source": "graphax src/graphax/examples/randoms.py:91-200", "known_answers": {"fwd": 632, "rev": 566},
       "args": args, "ops": ops, "outputs": outputs}
Path(__file__).with_name("random_g.json").write_text(json.dumps(out, indent=0))
print(len(args), len(ops), len(outputs))
