"""World-size-2 gloo test of the host-side multi-GPU logic (no GPU): the batch shards by image --
rank r of N encodes images r, r+N, ... -- and the job time is the max over ranks. No data-path
collective exists on this path; torch.distributed is plumbing only."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    import torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np, _libs
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    steps, w, h = 3, 48, 40
    # same shard rule as bench.py run_ours(): seed 1234 + 10*(step*world + rank)
    seeds = [1234 + 10 * (s * world + rank) for s in range(steps)]
    sums = [int(_libs.synth_image(w, h, sd).astype(np.int64).sum()) for sd in seeds]
    mine = torch.tensor([float(10 + rank)], dtype=torch.float64)   # pretend per-rank time
    dist.all_reduce(mine, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "seeds": seeds, "sums": sums})
    if rank == 0:
        print(json.dumps({"max_time": mine.item(), "shards": gathered}))
    dist.destroy_process_group()
''') % (ROOT, ROOT)


def test_batch_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["max_time"] == 11.0                      # max over ranks
    seeds = sorted(s for sh in res["shards"] for s in sh["seeds"])
    assert seeds == [1234 + 10 * k for k in range(6)]   # every image exactly once, none shared
    assert len({tuple(sh["sums"]) for sh in res["shards"]}) == 2   # ranks work on different images


def test_reference_arm_other_ranks_exit_without_work():
    """bench.py --impl reference: rank 0 alone runs; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


GROUP_WORKER = textwrap.dedent('''
    import os, sys, json
    import torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import __graft_entry__ as ge
    import test_quant_search_cpu as T
    gz = ge.load_package()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ev = T.make_eval("typical")
    # the product's own exchange path (torch.distributed all_gather; NCCL on the GPU box, gloo here)
    res = gz.QuantSearchSimulate(rank, world, gz.torch_allgather(dist, torch.device("cpu")), 0.971769, ev)
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        want, best = T.reference_search(ev, 0.971769)
        print(json.dumps({"res": gathered, "want": want, "best_q": best[0]}))
    dist.destroy_process_group()
''') % (ROOT, ROOT)


def test_candidate_sharded_quant_search_two_ranks_gloo(tmp_path):
    """SelectQuantMatrix candidates sharded over 2 ranks with the all-gather going through
    torch.distributed: both ranks replay the reference's visiting sequence in fewer rounds."""
    script = tmp_path / "group_worker.py"
    script.write_text(GROUP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    want = [tuple(v) for v in res["want"]]
    for r in res["res"]:
        got = [tuple(v) for v in r["visited"]]
        assert len(got) == len(want)
        assert all(g[0] == w[0] and abs(g[1] - w[1]) < 1e-9 and g[2] == w[2] and g[3] == w[3] for g, w in zip(got, want))
        assert r["best_q"] == res["best_q"]
    assert res["res"][0]["rounds"] == res["res"][1]["rounds"] < len(want)
    assert res["res"][0]["evaluated_here"] + res["res"][1]["evaluated_here"] == res["res"][0]["evaluated_total"]
