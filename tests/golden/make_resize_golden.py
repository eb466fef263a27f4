"""Golden vectors for `_resize_with_antialiasing`, produced by RUNNING the reference's own functions
(/root/reference/src/ctrlv/bbox_generator_baseline/utils/image_encoder.py:184-290, the copy that
src/ctrlv/utils/util.py:97-125 calls): the module cannot be imported here (it imports diffusers at
the top), so the five pure-torch function definitions are taken from its AST and executed as they
are.  Run in the build container (needs /root/reference); commits tests/golden/resize_antialias.pt."""
import ast
import os

import torch

SRC = "/root/reference/src/ctrlv/bbox_generator_baseline/utils/image_encoder.py"
NAMES = {"_resize_with_antialiasing", "_compute_padding", "_filter2d", "_gaussian", "_gaussian_blur2d"}
tree = ast.parse(open(SRC).read())
mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in NAMES], type_ignores=[])
ns = {"torch": torch}
exec(compile(mod, SRC, "exec"), ns)
assert NAMES <= set(ns)

g = torch.Generator().manual_seed(20240607)
cases = []
for shape, size in (((1, 3, 64, 96), (32, 32)), ((2, 3, 50, 70), (28, 28)), ((1, 3, 20, 24), (28, 28)),
                    ((1, 3, 96, 160), (16, 16))):
    x = torch.rand(shape, generator=g) * 2 - 1
    y = ns["_resize_with_antialiasing"](x, size)
    cases.append(dict(input=x, size=size, output=y))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_antialias.pt")
torch.save(cases, out)
print("wrote", out, [tuple(c["output"].shape) for c in cases], os.path.getsize(out), "bytes")
