"""Times the native train iteration at the reference's training shape: FLUX, bs 4 per GPU, latent 16x128x128
(S_img = 4096, S_txt = 512) — BASELINE.json configs[3]: 2 student + 8 teacher forwards, roll-out kernels, the full
adapter backward (per-block recompute), grad clip + AdamW + EMA and the bf16 write-back.
Usage: python tools/train_step_time.py [batch] [--profile] [--no-stash] [--kineto] [--bits8]
       --kineto: per-kernel time of ONE iteration from torch.profiler's CUPTI activity records (sees the library's kernels
                 too; concurrent timing, unlike ncu's serialised replay), aggregated by kernel name
   or: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_time.py [batch]
       (DDP: per-rank noise, ONE NCCL all-reduce of the flat gradient arena per iteration; time = max over ranks)"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from arcflow_b200 import _lib  # noqa: E402
from arcflow_b200.config import flux_dev  # noqa: E402
from arcflow_b200.model import ArcFluxEngineModel, FluxTeacherEngine  # noqa: E402
from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict, make_flux_teacher_extras  # noqa: E402
from arcflow_b200.train import ArcFlowTrainer, draw_rollout_randoms  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if args else 4
profile = "--profile" in sys.argv
kineto = "--kineto" in sys.argv
bits8 = "--bits8" in sys.argv
no_stash = "--no-stash" in sys.argv   # per-block recompute in the backward instead of the activation stash
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = flux_dev()
sd = make_flux_state_dict(cfg, 1234, dev)
student = ArcFluxEngineModel(sd, cfg, dev, consume_state_dict=True)
del sd
teacher = FluxTeacherEngine(student, make_flux_teacher_extras(cfg, 99, dev))
x, txt, pooled = make_flux_inputs(cfg, B, 1024, 1024, 512, 42 + rank, dev)
student.set_activation_stash(False if no_stash else "auto")
trainer = ArcFlowTrainer(student, teacher, state_bits=8 if bits8 else 32)
step = trainer.distill
g = torch.Generator().manual_seed(rank)
rands = [draw_rollout_randoms(B, 4, 16, g) for _ in range(2)]
lib = _lib.load()


def timed(fn, n):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = lib.afb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms_ = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
    return float(ms_.item()), (lib.afb_launch_count() - n0) // n, out


trainer.train_step(txt, pooled, (64, 64), x, rands, iteration=500)   # warm-up (also sizes every workspace)
fwd_ms, fwd_launches, (loss_f, _, _) = timed(lambda: step.forward(txt, pooled, (64, 64), x, rands, iteration=500), 2)
ms, launches, (loss, lv) = timed(lambda: trainer.train_step(txt, pooled, (64, 64), x, rands, iteration=500), 2)
fwd_flops = (2 * 78.77e12 + 8 * 74.36e12) * B
out = dict(what="train iteration (fwd + bwd + optimizer)", n_gpus=world, batch=B, ms=ms, forward_only_ms=fwd_ms, backward_optim_ms=ms - fwd_ms,
           samples_per_s=B * world / (ms / 1e3), loss=loss, fwd_tflops=fwd_flops / (fwd_ms * 1e9), launches=launches,
           fwd_launches=fwd_launches, mem_gb=torch.cuda.max_memory_allocated() / 2**30,
           activation_stash=bool(student.activation_stash),
           device_mem_used_gb=(torch.cuda.mem_get_info(dev)[1] - torch.cuda.mem_get_info(dev)[0]) / 2**30,
           log_vars={k: (float(v) if isinstance(v, (int, float)) else v) for k, v in lv.items()})
if profile:
    student.set_profiling(True)
    trainer.train_step(txt, pooled, (64, 64), x, rands, iteration=500)
    out["student_profile"] = student.read_profile()
    student.set_profiling(False)
if kineto and rank == 0:
    import re
    from collections import defaultdict
    from torch.profiler import ProfilerActivity, profile as tprofile
    with tprofile(activities=[ProfilerActivity.CUDA]) as prof:
        trainer.train_step(txt, pooled, (64, 64), x, rands, iteration=500)
        torch.cuda.synchronize()
    agg, cnt = defaultdict(float), defaultdict(int)
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.device_time > 0:
            name = re.sub(r"\(anonymous namespace\)::|afb::|<unnamed>::|void ", "", ev.name)
            name = re.sub(r"\(.*$", "", name)[:70]
            agg[name] += ev.device_time / 1e3
            cnt[name] += 1
    total = sum(agg.values())
    out["kineto"] = dict(total_kernel_ms=total, kernels=[dict(name=k, ms=round(v, 3), share=round(v / total, 4), launches=cnt[k])
                                                         for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:40]])
if world > 1:   # DDP invariant: every rank holds the same parameters after the step
    chk = trainer.opt.params.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["params_identical_across_ranks"] = bool((lo == hi).item())
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
