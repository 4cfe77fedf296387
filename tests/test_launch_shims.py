"""The launcher's environment glue (turboae_b200/launch.py::_modernise), run in a SUBPROCESS against the reference staged in
baseline/_ref (git-ignored; skipped when absent): without it the reference's own classical turbo encoder and its Lookahead
optimizer fail on Python 3 / torch >= 2 before any of this package's code is reached."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

REF = os.path.join(ROOT, "baseline", "_ref")

CODE = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from turboae_b200 import launch
launch._modernise()
from commpy.channelcoding.interleavers import RandInterlv
r = RandInterlv(12, 3)
x = np.arange(100, 112)
y = r.interlv(x)
assert y.shape == (12,) and np.array_equal(y, x[r.p_array]), y          # Python 2's array(map(...)) gave a 0-d object array
assert np.array_equal(r.deinterlv(y), x)
import optimizers
p = torch.nn.Parameter(torch.ones(3))
opt = optimizers.Lookahead(torch.optim.SGD([p], lr=0.1))
p.grad = torch.ones(3)
opt.zero_grad()                                                          # torch >= 2 reads self.defaults here
assert p.grad is None or float(p.grad.abs().sum()) == 0.0
p.grad = torch.ones(3)
opt.step()
assert abs(float(p.detach()[0]) - 0.9) < 1e-6                            # SGD step to 0.9; the first sync starts the slow weights there
print("ok")
"""


def test_launcher_shims_make_the_reference_turbo_encoder_and_lookahead_run():
    if not os.path.isfile(os.path.join(REF, "main.py")):
        pytest.skip("baseline/_ref not staged (python scripts/stage_reference.py)")
    r = subprocess.run([sys.executable, "-c", CODE % (REF, ROOT)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


FUSED_CODE = r"""
import sys, torch
sys.path.insert(0, %r)
from turboae_b200 import launch
launch.fused_adam_default(); launch.fused_adam_default()                  # idempotent
m = torch.nn.Linear(3, 3)
# CPU parameters: torch's own choice stays (fused needs CUDA); the reference's filter(...) generator is materialised once
o = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=1e-3)
assert o.defaults.get("fused") is None and len(o.param_groups[0]["params"]) == 2 and o.defaults["lr"] == 1e-3
o = torch.optim.Adam(m.parameters(), 1e-3, foreach=True)                  # an explicit choice is respected
assert o.defaults.get("fused") is None and o.defaults["foreach"] is True
# what the wrapper decides for device parameters, without a device: stand-in leaves that report is_cuda
class Fake(torch.Tensor):
    is_cuda = True
seen = {}
_orig = torch.optim.Adam.__init__.__closure__[0].cell_contents
def spy(self, params, *a, **k):
    seen.update(k)
    raise RuntimeError("stop")
import types
for c in torch.optim.Adam.__init__.__closure__:
    if c.cell_contents is _orig:
        c.cell_contents = spy
try:
    torch.optim.Adam([torch.zeros(2).as_subclass(Fake)], lr=1e-3)
except RuntimeError:
    pass
assert seen.get("fused") is True, seen
print("ok")
"""


def test_launcher_fused_adam_default():
    """launch.fused_adam_default: fused=True only when every parameter is a CUDA floating tensor and the caller chose nothing."""
    out = subprocess.run([sys.executable, "-c", FUSED_CODE % ROOT], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
