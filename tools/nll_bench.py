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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(lib.theta_solve_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), stt.data_ptr(), None))
    e1.record()
    torch.cuda.synchronize()
    print("solve_tiles: %.3f ms per launch, %.3e candidates/s" % (e0.elapsed_time(e1) / 5, cnt / (e0.elapsed_time(e1) / 5 * 1e-3)))
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

def n3():
    """n = 3 rows (BASELINE configs[3] shape, m=150, k=5): 2^22 real candidates of the search, mu from the solver."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O
    from theta_b200 import _lib
    _lib.init(0)
    torch.cuda.set_device(0)
    P3 = O.synthetic_problem_n3(150, 5, 1, 12, False)
    plan = _lib.Plan(3, P3["r"], P3["rN"], P3["lower"], P3["upper"])
    n_items = plan.prepare(512)
    paths = plan.enumerate_n3(0, min(n_items, 4000))
    mu, mh, nll_ref, st, _ = plan.solve_rows(paths)
    ok = (st < 8) & (st != _lib.ST_DEGENERATE)
    paths, mu, nll_ref = paths[ok], mu[ok], nll_ref[ok]
    reps_n = (1 << 22) // len(paths) + 1
    cnt = 1 << 22
    t = np.zeros((len(paths), plan.pitch), dtype=np.uint8)
    t[:, :plan.m] = paths
    tiles = torch.from_numpy(np.tile(t, (reps_n, 1))[:cnt].copy()).cuda()
    mus = torch.from_numpy(np.tile(mu, (reps_n, 1))[:cnt].copy()).cuda()
    ref = torch.from_numpy(np.tile(nll_ref, reps_n)[:cnt].copy()).cuda()
    out = torch.zeros(cnt, dtype=torch.float64, device="cuda")
    lib = _lib.load()
    for _ in range(3):
        _lib.check(lib.theta_nll_rows_n3_dev(plan.handle, tiles.data_ptr(), plan.pitch, cnt, mus.data_ptr(), out.data_ptr(), None))
    torch.cuda.synchronize()
    err = float(((out - ref).abs() / ref.abs()).max())
    reps = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        _lib.check(lib.theta_nll_rows_n3_dev(plan.handle, tiles.data_ptr(), plan.pitch, cnt, mus.data_ptr(), out.data_ptr(), None))
        ev[i + 1].record()
    torch.cuda.synchronize()
    tms = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(reps)) / reps
    runs = float((np.diff(paths.astype(np.int16), axis=1) != 0).sum(axis=1).mean()) + 1
    print("n3 m=%d pitch=%d cnt=%d distinct=%d mean runs/row %.1f  mean %.4f ms  %.0f GB/s (m+24+8 B/candidate)  maxrel %.2e" %
          (plan.m, plan.pitch, cnt, len(paths), runs, tms, cnt * (plan.m + 32) / (tms * 1e-3) / 1e9, err))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--n3":
        n3()
    elif len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        wls = sys.argv[1:] or ["cfg2"]
        for wl in wls:
            for v in ("2", "4", "5", "6"):
                env = dict(os.environ, THETA_NLL_VARIANT=v)
                subprocess.run([sys.executable, os.path.abspath(__file__), "--one", wl], env=env)
