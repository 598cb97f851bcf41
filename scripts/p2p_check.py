#!/usr/bin/env python
"""Multi-process check of the peer-memory sharded path (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/p2p_check.py

Every rank trains a P2PShardedDLRM for a few steps on its own batches; rank 0 also runs the same global problem UNSHARDED
and compares the reassembled table and the probabilities (recommender_b200/p2p_selfcheck.py).  fp32 towers: the sharded
and the unsharded arithmetic differ only in the order the replicas' dense gradients are added (tolerance 2e-6); bf16
towers: a rounding boundary crossed by such a difference moves an Adam row update by a few 1e-6 (tolerance 1e-4)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from recommender_b200 import p2p_selfcheck
    ok = True
    unequal = [5000] * 20 + [3, 14, 63, 155, 976, 2208]          # small tables beside sharded ones (BASELINE config 3 in miniature)
    cases = [(None, 2e-6, {}), (torch.bfloat16, 1e-4, {}),
             (None, 2e-6, dict(table_rows=unequal)), (None, 2e-6, dict(table_rows=unequal, replicate_rows_upto=4096)),
             (torch.bfloat16, 1e-4, dict(table_rows=unequal, replicate_rows_upto=4096))]
    for dtype, tol, kw in cases:
        res = p2p_selfcheck.run(dev, compute_dtype=dtype, **kw)
        if rank == 0:
            good = res["max_abs_table_diff"] <= tol and res["rows_moved"] > 0 and res["replica_max_abs_diff"] == 0.0
            good &= res["replicated_tables"] == (6 if kw.get("replicate_rows_upto") else 0)
            ok &= good
            print("p2p_check:", json.dumps(res), "OK" if good else "MISMATCH", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
