"""world_size-2 gloo test of bench.py's multi-rank plumbing (barrier, MAX over ranks, id broadcast, per-rank
shards of the synthetic workload).  CPU only; the GPU path uses the same helpers with the nccl backend."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "a1-qp-mpc-controller_b200"))
import numpy as np
import bench, a1mpc
rank, world, local, dist = bench.dist_setup(2)
assert world == 2 and dist is not None
bench.dist_barrier(dist)
m = bench.dist_max(dist, 10.0 + rank)          # max over ranks of the per-rank device time
assert m == 11.0, m
uid = bench.dist_bcast_bytes(dist, b"x" * 128 if rank == 0 else None, rank)
assert uid == b"x" * 128
st = a1mpc.gen_states(256, 2, stream=rank * 1000003)     # each rank owns an independent slice
import torch
t = torch.tensor(st["x0"][:, :4].copy())
out = [torch.zeros_like(t) for _ in range(2)]
dist.all_gather(out, t)
assert not torch.equal(out[0], out[1])                   # shards differ
assert torch.equal(out[rank], t)
# the collect set-up must take the same branch on every rank and never leave a rank waiting in a collective: rank 1 cannot create
# its peer buffer, rank 0 cannot produce an NCCL id -> both end with "unavailable", no hang
class FakeEng:
    def peer_gather_create(self, n, r, B):
        if r == 1:
            raise RuntimeError("no peer access (test)")
        return b"h" * 64
    def peer_gather_connect(self, hs): pass
    def peer_gather_destroy(self): pass
    def peer_gather_buffer(self): return None
    def nccl_init(self, *a): raise RuntimeError("nccl_init must not be reached")
class FakeA1:
    @staticmethod
    def nccl_unique_id(): raise RuntimeError("no NCCL (test)")
desc, fn, err, buf = bench.setup_collect(FakeA1, FakeEng(), dist, 2, rank, 64, "auto")
assert desc == "unavailable" and fn is None and "no NCCL (test)" in err, (desc, err)
desc, fn, err, buf = bench.setup_collect(FakeA1, FakeEng(), dist, 2, rank, 64, "peer")
assert desc == "unavailable" and fn is None
assert bench.dist_allgather_obj(dist, rank * 7, 2) == [0, 7]
bench.dist_barrier(dist)
sys.stdout.write("RANK_OK " + str(rank) + "\n"); sys.stdout.flush()    # one write per rank: the two ranks share a pipe
'''


def test_two_rank_plumbing(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % (ROOT, ROOT))
    import socket
    with socket.socket() as sk:      # a free rendezvous port (back-to-back runs would otherwise collide in TIME_WAIT)
        sk.bind(("127.0.0.1", 0))
        port = str(sk.getsockname()[1])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", port, str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "RANK_OK 0" in r.stdout and "RANK_OK 1" in r.stdout
