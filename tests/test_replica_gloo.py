"""CPU, world_size 2 over gloo: the replica-parallel control plane used by bench.py --gpus N."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import json, os, sys
    sys.path.insert(0, %r)
    from nunchaku_b200 import replica

    rank, ws, _ = replica.world()
    replica.init("gloo")
    job = replica.broadcast_job({"images": list(range(7)), "seed": 1234} if rank == 0 else None)
    mine = replica.stripe(job["images"], rank, ws)
    slow = replica.max_over_ranks(10.0 + rank)           # rank 1 is the slow one
    res = replica.gather_results({"rank": rank, "units": mine, "seed": job["seed"]})
    if rank == 0:
        print("RESULT " + json.dumps({"gathered": res, "max_time": slow, "ws": ws}), flush=True)
    replica.shutdown()
    """
) % ROOT


def test_two_rank_gloo_roundtrip(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=env)
    assert pr.returncode == 0, pr.stdout[-2000:] + pr.stderr[-2000:]
    line = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")][0]
    import json

    r = json.loads(line[len("RESULT "):])
    assert r["ws"] == 2 and r["max_time"] == 11.0
    units = sorted(u for g in r["gathered"] for u in g["units"])
    assert units == list(range(7))                          # every unit exactly once
    assert [g["units"] for g in sorted(r["gathered"], key=lambda g: g["rank"])] == [[0, 2, 4, 6], [1, 3, 5]]
    assert all(g["seed"] == 1234 for g in r["gathered"])    # broadcast reached every rank


def test_single_process_defaults():
    from nunchaku_b200 import replica

    assert replica.world() == (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))
    assert replica.stripe([1, 2, 3], 0, 1) == [1, 2, 3]
    assert replica.broadcast_job({"a": 1}) == {"a": 1}
    assert replica.gather_results(5) == [5]
    assert replica.max_over_ranks(3.5) == 3.5
