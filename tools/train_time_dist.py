"""Per-rank training-step time under torchrun with a FIXED per-GPU batch (what every GPU of the 8-GPU strong-scaling run
executes: 64 videos), CUDA graph + bucketed NCCL all-reduce:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/train_time_dist.py [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from video2music_b200 import VideoMusicTransformer, synthetic as syn
from video2music_b200.trainer import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = VideoMusicTransformer(total_vf_dim=syn.vf_dim(0), rpr=True, dropout=0.2)
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict(syn.fill_like_reference_init(shapes, seed=1), strict=False)
m = m.to(dev).train().set_compute_dtype(torch.bfloat16)
tr = Trainer(m, use_graph=True)
b = {k: v.to(dev) for k, v in syn.make_inputs(B, 1234 + rank, 299, 300, 0).items()}
for _ in range(5):
    loss = tr.train_step(b)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.train_step(b)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("world %d, %d videos per GPU: %.2f ms/step (max over ranks), loss %.4f, env NCCL_MAX_NCHANNELS=%s NCCL_MAX_CTAS=%s V2M_RESERVE_SMS=%s" % (
        world, B, float(t[0]), float(loss), os.environ.get("NCCL_MAX_NCHANNELS"), os.environ.get("NCCL_MAX_CTAS"), os.environ.get("V2M_RESERVE_SMS")))
if world > 1:
    dist.destroy_process_group()
