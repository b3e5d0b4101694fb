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
for it in range(3):
    out = head(None, None, b.planes1, b.planes2, b.app1, b.app2, matching_net=match, initial_pose=ip, result_exchange=ex)
    rows = ex.finish().clone()
    ref = gather_results(out[5]["pose"].contiguous(), world * B)
    torch.cuda.synchronize()
    same = bool(torch.equal(rows, ref))
    ok = ok and same
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("FUSED_EXCHANGE_OK" if int(flag.item()) == 1 else "FUSED_EXCHANGE_MISMATCH", "world", world, "rows", tuple(rows.shape))
dist.destroy_process_group()
