"""The N>1 path on CPU: two gloo ranks shard the batch, run the hooked model (oracle quantizers injected - no GPU
here), and combine the four validation counters with the single all-reduce the path has."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from cnn_quantization_b200 import pipeline
from oracle import fq_oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.set_num_threads(2)
model, qm = pipeline.build_quantized_model("resnet18_w4a4", "cpu", quantizer_factory=O.oracle_int_quantizer)
x, t = pipeline.synthetic_batch(4, seed=7, hw=64)            # the global batch, identical on every rank
shard = slice(rank * 4 // world, (rank + 1) * 4 // world)       # rank r takes images [r*B/G, (r+1)*B/G)
total = pipeline.validate(model, [(x[shard], t[shard])], "cpu")
local = total.clone()
loss, top1, top5, n = pipeline.reduce_metrics(total)
# a replica run on the same shard alone gives the same local counters (statistics are per shard, SURVEY 8e)
if rank == 0:
    print(json.dumps({"n": n, "loss": loss, "top1": top1, "top5": top5, "local_n": float(local[3])}))
dist.destroy_process_group()
""" % ROOT


def test_two_rank_gloo_sharding_and_single_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["n"] == 4 and out["local_n"] == 2.0
    assert 0 <= out["top1"] <= out["top5"] <= 100
