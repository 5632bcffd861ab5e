"""Streaming NLL kernel: GB/s of each launch variant (THETA_NLL_VARIANT is read once per process, so
each variant runs in its own subprocess).  Development aid; prints achieved algorithmic GB/s."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def one(wl):
    import torch
    import bench
    from theta_b200 import _lib
    _lib.init(0)
    torch.cuda.set_device(0)
    P = bench.problem(wl)
    plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
    lib = _lib.load()
    cnt = min(plan.count(), 4 * 1024 * 1024)
    pitch = plan.pitch
    tiles = torch.empty(cnt * pitch, dtype=torch.int8, device="cuda")
    mu = torch.empty(cnt, dtype=torch.float64, device="cuda")
    nll = torch.zeros(cnt, dtype=torch.float64, device="cuda")
    stt = torch.empty(cnt, dtype=torch.int32, device="cuda")
    _lib.check(lib.theta_enumerate_n2_dev(plan.handle, plan.count() // 3, cnt, tiles.data_ptr(), pitch, None))
    _lib.check(lib.theta_solve_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), stt.data_ptr(), None))
    torch.cuda.synchronize()
    ref = nll.clone()
    nll.zero_()
    for _ in range(3):
        _lib.check(lib.theta_nll_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), None))
    torch.cuda.synchronize()
    ok = stt < 8
    err = float(((nll[ok] - ref[ok]).abs() / ref[ok].abs()).max())
    reps = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        _lib.check(lib.theta_nll_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), None))
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    t = sum(ts) / reps
    gbs = cnt * (plan.m + 16) / (t * 1e-3) / 1e9
    print("variant %s wl %s m=%d pitch=%d cnt=%d  mean %.4f ms  min %.4f ms  %.0f GB/s (mean)  %.0f GB/s (best)  maxrel %.2e" %
          (os.environ.get("THETA_NLL_VARIANT", "default"), wl, plan.m, pitch, cnt, t, ts[0], gbs, cnt * (plan.m + 16) / (ts[0] * 1e-3) / 1e9, err))

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        wls = sys.argv[1:] or ["cfg2"]
        for wl in wls:
            for v in ("2", "4", "5", "6"):
                env = dict(os.environ, THETA_NLL_VARIANT=v)
                subprocess.run([sys.executable, os.path.abspath(__file__), "--one", wl], env=env)
