#!/usr/bin/env python
"""
Timeline of the screening pass at config C3 (N=4096, d=10): per window, when the fp32 cross-covariance of window w+1 (side
stream) and the low-digit tensor-core product of window w run, how long each takes with and without the stream overlap, for
2, 3 and 4 screening digits.

    python tools/screen_trace.py [candidates] [out.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    import torch

    from pygpso_b200 import backend

    M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_100_000
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "screen_trace.json")
    N, d = 4096, 10
    X, y = bench.synthetic_training(N, d)
    theta = bench.fixed_theta(d)
    cuda = backend.CudaBackend(device=0)
    sess = cuda.open_session("Matern52", 1, True)
    sess.set_data(X, y)
    xc = torch.empty((M, d), dtype=torch.float64)
    bench.fill_candidates(xc.numpy(), 0, M, 10_000_000, d)
    xc = xc.cuda()
    stream = torch.cuda.current_stream().cuda_stream
    vs = bench.VARSIGMA
    report = {"M": M, "N": N, "d": d, "peaks": cuda.probe_peaks()}

    def timed(steps):
        sess.ucb_argmax_dev(xc.data_ptr(), M, vs, stream)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        prod = 0.0
        for _ in range(steps):
            res = sess.ucb_argmax_dev(xc.data_ptr(), M, vs, stream)
            prod += sess.screen_info()["screen_product_ms"]
        ev1.record()
        torch.cuda.synchronize()
        info = sess.screen_info()
        ms = ev0.elapsed_time(ev1) / steps
        return {"ms_per_step": ms, "screen_product_ms_per_step": prod / steps, "cand_per_s": M / ms * 1e3, "argmax": int(res[0]),
                "path": info["path"], "survivors": info["survivors"], "windows": info["screen_windows"], "E": info["error_bound"],
                "max_dev": info["max_observed_deviation"]}

    for digits in (3, 2, 4, 6) if len(sys.argv) < 4 else ((6,) if sys.argv[3] == "2f" else (3, 6)):  # 6 = two digits, all four digit pairs
        sess.set_screen_mode(digits)
        sess.factorize(theta)
        rec = {}
        sess.set_overlap(True)
        rec["overlap"] = timed(3)
        sess.set_overlap(False)
        rec["no_overlap"] = timed(3)
        sess.set_overlap(True)
        sess.set_profile(2)
        sess.ucb_argmax_dev(xc.data_ptr(), M, vs, stream)
        tr = sess.trace()
        nwin = sess.screen_info()["screen_windows"]
        sess.set_profile(0)
        rows = {}
        seen = {}
        for tag, w, ms in tr:
            key = (int(w), int(tag))
            if key in seen:  # marks of the refine pass reuse window numbers: keep the screening pass (first occurrence)
                continue
            seen[key] = True
            rows.setdefault(int(w), {})[int(tag)] = ms
        windows = []
        for w in sorted(rows):
            if w >= nwin:
                continue
            r = rows[w]
            item = {"w": w, "xcov_start": r.get(1), "xcov_end": r.get(2), "prod_start": r.get(3), "prod_end": r.get(4), "fin_end": r.get(5),
                    "xcov_ms": r[2] - r[1], "prod_ms": r[4] - r[3], "fin_ms": r[5] - r[4]}
            if w - 1 in rows:
                item["gap_after_prev_fin_ms"] = r[3] - rows[w - 1][5]
                item["prod_start_minus_xcov_end_ms"] = r[3] - r[2]
            windows.append(item)
        rec["timeline"] = windows
        inner = windows[1:-1] if len(windows) > 2 else windows
        rec["summary"] = {k: float(np.mean([x[k] for x in inner if k in x])) for k in
                          ("xcov_ms", "prod_ms", "fin_ms", "gap_after_prev_fin_ms", "prod_start_minus_xcov_end_ms")}
        report[f"digits{digits}"] = rec
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(report, fh, indent=1)
    for k, v in report.items():
        if k.startswith("digits"):
            print(k, json.dumps({"overlap": v["overlap"], "no_overlap": v["no_overlap"], "summary": v["summary"]}))
    sess.close()


if __name__ == "__main__":
    main()
