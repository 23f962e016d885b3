"""Run every `-m gpu` test in its own process (a trapped kernel poisons the CUDA context, so
isolation keeps one bad kernel from hiding the rest) and write logs under gpurun_out/.

    python tools/gpu_run_tests.py [-k EXPR] [--timeout 240] [--tag NAME]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--tag", default="tests")
    ap.add_argument("--files", nargs="*", default=["tests"])
    args = ap.parse_args()
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    cmd = [sys.executable, "-m", "pytest", *args.files, "-m", "gpu", "--collect-only", "-q"]
    if args.k:
        cmd += ["-k", args.k]
    col = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    ids = [l.strip() for l in col.stdout.splitlines() if "::" in l]
    print(f"collected {len(ids)} gpu tests", flush=True)
    results = {}
    log_path = os.path.join(out_dir, f"{args.tag}.log")
    with open(log_path, "w") as log:
        for tid in ids:
            t0 = time.time()
            try:
                pr = subprocess.run([sys.executable, "-m", "pytest", tid, "-q", "-x", "--no-header", "-p", "no:cacheprovider"],
                                    cwd=ROOT, capture_output=True, text=True, timeout=args.timeout)
                status = "pass" if pr.returncode == 0 else "FAIL"
                text = pr.stdout[-6000:] + pr.stderr[-3000:]
            except subprocess.TimeoutExpired as e:
                status = "TIMEOUT"
                text = (e.stdout or b"").decode(errors="replace")[-3000:] if isinstance(e.stdout, bytes) else str(e.stdout)[-3000:]
            dt = time.time() - t0
            results[tid] = {"status": status, "seconds": round(dt, 1)}
            print(f"{status:8s} {dt:6.1f}s {tid}", flush=True)
            log.write(f"===== {status} {tid} ({dt:.1f}s)\n")
            if status != "pass":
                log.write(text + "\n")
            log.flush()
    with open(os.path.join(out_dir, f"{args.tag}.json"), "w") as f:
        json.dump(results, f, indent=1)
    n_bad = sum(1 for r in results.values() if r["status"] != "pass")
    print(f"{len(ids) - n_bad}/{len(ids)} passed; log: {log_path}")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
