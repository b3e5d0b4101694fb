"""torchrun --nproc-per-node N scripts/check_fused_exchange.py — the fused NVLink result exchange (selection kernel
stores rows into every rank's symmetric buffer) must equal one NCCL all-gather of the local result rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from nopesac_b200 import synthetic
from nopesac_b200.dist import FusedResultExchange, gather_results
from tests import util

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, P, NQ = 6, 16, 50
head, match, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
b = synthetic.make_batch(rank * B, B, P).to(dev)
poses = [util.initial_pose_for(200 + rank * B + i) for i in range(B)]
ip = (torch.cat([p[0] for p in poses]).to(dev), torch.cat([p[1] for p in poses]).to(dev))
ex = FusedResultExchange(B, dev)
ok = True
prev_rows = prev_ref = None
for it in range(6):
    # different inputs every step (so a stale slot cannot pass), a deliberately late rank on odd steps
    bb = synthetic.make_batch(1000 * it + rank * B, B, P).to(dev)
    if it % 2 == 1 and rank == world - 1:
        torch.cuda._sleep(20_000_000)          # ~10 ms of device time: the other ranks run ahead into the next step
    out = head(None, None, bb.planes1, bb.planes2, bb.app1, bb.app2, matching_net=match, initial_pose=ip, result_exchange=ex)
    rows = ex.finish()                          # NOT cloned: the slot must stay valid through the next step (double buffer)
    ref = gather_results(out[5]["pose"].contiguous(), world * B)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(rows, ref))
    if prev_rows is not None:                   # the previous step's slot is untouched by this step's peer stores
        ok = ok and bool(torch.equal(prev_rows, prev_ref))
    prev_rows, prev_ref = rows, ref
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("FUSED_EXCHANGE_OK" if int(flag.item()) == 1 else "FUSED_EXCHANGE_MISMATCH", "world", world, "rows", tuple(rows.shape))
dist.destroy_process_group()
