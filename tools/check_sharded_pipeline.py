"""Run under `torchrun --nproc-per-node N`: every rank makes the SAME `pipe(...)` call (NCCL); the pipeline shards the batch,
denoises its images and all-gathers the final latents. Rank 0 also runs the whole batch alone with batch-parallel switched
off and writes a JSON verdict: the sharded result must be bit-identical (images are independent units and the kernels are
batch-invariant), whatever the rank count. Usage: torchrun ... tools/check_sharded_pipeline.py out.json [batch]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from arcflow_b200.config import flux_tiny  # noqa: E402
from arcflow_b200.model import ArcFluxEngineModel  # noqa: E402
from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict  # noqa: E402
from lakonlab.parallel import init_from_env  # noqa: E402
from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline  # noqa: E402

out_path = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank, world, dev = init_from_env()
cfg = flux_tiny(2, 2, 2)
sd = make_flux_state_dict(cfg, seed=1234, device="cpu")
x, txt, pooled = make_flux_inputs(cfg, batch, 64, 64, txt_len=32, seed=9)
pipe = ArcFluxPipeline(transformer=ArcFluxEngineModel(sd, cfg, device=dev))
kw = dict(prompt_embeds=txt, pooled_prompt_embeds=pooled, height=64, width=64, num_inference_steps=2, timestep_ratio=1.0,
          output_type="latent")
sharded = pipe(latents=x, **kw).images                                   # collective: all ranks
gen = pipe(generator=torch.Generator(dev).manual_seed(7), **kw).images   # noise drawn inside the call
res = dict(world=world, batch=batch)
gathered = [torch.empty_like(sharded) for _ in range(world)]
dist.all_gather(gathered, sharded.contiguous())
res["all_ranks_agree"] = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    pipe.enable_batch_parallel(False)
    alone = pipe(latents=x, **kw).images
    alone_gen = pipe(generator=torch.Generator(dev).manual_seed(7), **kw).images
    res.update(shape=list(sharded.shape), equals_single_rank=bool(torch.equal(alone, sharded)),
               generator_path_equals_single_rank=bool(torch.equal(alone_gen, gen)), finite=bool(torch.isfinite(sharded).all()))
    with open(out_path, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
