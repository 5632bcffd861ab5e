"""Short driver for ncu captures: a couple of launches of each hot kernel (no timing claims)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from theta_b200 import _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
_lib.init(0)
torch.cuda.set_device(0)
P = bench.headline(1) if wl == "headline" else bench.problem(wl)  # "headline": the window bench.py times at N=1
plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
N = plan.count()
if P["window"]:
    N = min(N, P["window"])
if what in ("all", "search"):
    for _ in range(2):
        g = plan.pass1(0, min(N, 1 << 26))
        plan.pass2(g)
if what in ("all", "tiles"):
    lib = _lib.load()
    cnt = min(plan.count(), 2 * 1024 * 1024)
    pitch = plan.pitch
    tiles = torch.empty(cnt * pitch, dtype=torch.int8, device="cuda")
    mu = torch.empty(cnt, dtype=torch.float64, device="cuda")
    nll = torch.empty(cnt, dtype=torch.float64, device="cuda")
    stt = torch.empty(cnt, dtype=torch.int32, device="cuda")
    _lib.check(lib.theta_enumerate_n2_dev(plan.handle, plan.count() // 3, cnt, tiles.data_ptr(), pitch, None))
    _lib.check(lib.theta_solve_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), stt.data_ptr(), None))
    for _ in range(2):
        _lib.check(lib.theta_nll_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), None))
    torch.cuda.synchronize()
print("done")
