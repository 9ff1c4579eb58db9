#!/usr/bin/env python
"""
Timeline of the scoring pipeline at config C3 (N=4096, d=10): when do the cross-covariance kernel of window w+1 (side
stream, FP64 CUDA cores) and the tensor-core product of window w actually run, how long is each under overlap, and what do
SM clock and board power look like with and without the overlap.

    python tools/overlap_trace.py [candidates] [out.json]
"""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


class PowerSampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows = []
        self.proc = None

    def run(self):
        self.proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                      "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            try:
                a, b = line.split(",")
                self.rows.append((float(a), float(b)))
            except ValueError:
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        if not self.rows:
            return None
        r = np.array(self.rows)
        return {"sm_mhz_median": float(np.median(r[:, 0])), "power_w_median": float(np.median(r[:, 1])),
                "power_w_max": float(r[:, 1].max()), "samples": len(r)}


def main():
    import torch

    from pygpso_b200 import backend

    M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_200_000
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "overlap_trace.json")
    N, d = 4096, 10
    X, y = bench.synthetic_training(N, d)
    theta = bench.fixed_theta(d)
    cuda = backend.CudaBackend(device=0)
    sess = cuda.open_session("Matern52", 1, True)
    sess.set_data(X, y)
    sess.factorize(theta)
    xc = torch.from_numpy(np.random.default_rng([bench.SEED, 0]).random((M, d))).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    vs = bench.VARSIGMA
    report = {"M": M, "N": N, "d": d, "engine": sess.predict_info()}

    def timed(steps, label):
        sess.ucb_argmax_dev(xc.data_ptr(), M, vs, stream)
        torch.cuda.synchronize()
        sampler = PowerSampler()
        sampler.start()
        time.sleep(0.3)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        prod = 0.0
        for _ in range(steps):
            res = sess.ucb_argmax_dev(xc.data_ptr(), M, vs, stream)
            prod += sess.last_timing_ms()[2]
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        report[label] = {"ms_per_step": ms, "product_ms_per_step": prod / steps, "windows": sess.last_windows(),
                         "cand_per_s": M / ms * 1e3, "gpu": sampler.stop(), "argmax": int(res[0])}

    steps = max(3, int(6e6 // M))
    sess.set_overlap(True)
    timed(steps, "overlap")
    sess.set_overlap(False)
    timed(steps, "no_overlap")
    sess.set_overlap(True)
    # persisting-L2 window over the A digits off / on again (same process, same GPU)
    sess.set_l2_window(False)
    sess.factorize(theta)
    timed(steps, "l2_window_off")
    sess.set_l2_window(True)
    sess.factorize(theta)
    timed(steps, "l2_window_on_again")

    # per-stage sums with the stages run back to back (no overlap)
    sess.set_profile(1)
    sess.ucb_argmax_dev(xc.data_ptr(), M, vs, stream)
    t = sess.last_timing_ms()
    nw = sess.last_windows()
    report["serialised_ms_per_window"] = {"crosscov": t[1] / nw, "product": t[2] / nw, "finalize": t[3] / nw}
    sess.set_profile(0)

    # timeline of one overlapped step
    sess.set_profile(2)
    sess.ucb_argmax_dev(xc.data_ptr(), M, vs, stream)
    tr = sess.trace()
    sess.set_profile(0)
    rows = {}
    for tag, w, ms in tr:
        rows.setdefault(int(w), {})[int(tag)] = ms
    windows = []
    for w in sorted(rows):
        r = rows[w]
        rec = {"w": w, "xcov_start": r.get(1), "xcov_end": r.get(2), "prod_start": r.get(3), "prod_end": r.get(4), "fin_end": r.get(5)}
        rec["xcov_ms"] = r[2] - r[1]
        rec["prod_ms"] = r[4] - r[3]
        rec["fin_ms"] = r[5] - r[4]
        if w - 1 in rows:
            rec["gap_after_prev_fin_ms"] = r[3] - rows[w - 1][5]
            rec["prod_start_minus_xcov_end_ms"] = r[3] - r[2]
        windows.append(rec)
    report["timeline"] = windows
    inner = windows[1:-1] if len(windows) > 2 else windows
    report["summary"] = {k: float(np.mean([x[k] for x in inner if k in x])) for k in
                         ("xcov_ms", "prod_ms", "fin_ms", "gap_after_prev_fin_ms", "prod_start_minus_xcov_end_ms")}
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(report, fh, indent=1)
    print(json.dumps({k: report[k] for k in ("overlap", "no_overlap", "l2_window_off", "l2_window_on_again", "serialised_ms_per_window", "summary")}, indent=1))
    sess.close()


if __name__ == "__main__":
    main()
