"""N > 1 host logic on CPU: world_size-2 gloo run of the batch-parallel shard + sampler-boundary all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lakonlab.parallel.batch_parallel import gather_latents, shard_batch, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 6, dtype=torch.float32).reshape(total, 3, 2)
        local = shard_batch(full)                      # this rank's images
        local = local * 2.0 + 1.0                      # stand-in for the per-image denoise (no cross-image dependence)
        out = gather_latents(local, total)
        q.put((rank, torch.equal(out, full * 2.0 + 1.0), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5, 2])
def test_shard_and_gather_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, f"rank {rank} gathered a wrong tensor"
        assert shape == (total, 3, 2)


def test_shard_bounds_cover_everything():
    for total in (1, 2, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    x = torch.randn(3, 4, 64)
    assert shard_batch(x, 0, 1) is not None and torch.equal(shard_batch(x, 0, 1), x)
    assert torch.equal(gather_latents(x, 3), x)


def _opt_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arcflow_b200.optim import FlatAdamW
        o = FlatAdamW.__new__(FlatAdamW)          # host-side logic only: the arena all-reduce (no CUDA on this box)
        o.grads = torch.full((10,), float(rank + 1))
        o.all_reduce_grads()
        q.put((rank, o.grads.tolist()))
    finally:
        dist.destroy_process_group()


def test_gradient_arena_all_reduce_is_a_mean_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_opt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, g in res:
        assert g == [1.5] * 10


# ------------------------------------------------------------------------------------------------
# lakonlab.apis.train_model under world_size 2 (gloo): rank-disjoint data shards, all ranks step together, only rank 0
# writes checkpoints — the host protocol of `torchrun ... train.py --launcher pytorch` with the engine replaced by a stub.
# ------------------------------------------------------------------------------------------------
class _StubOpt:
    def __init__(self):
        self.v = torch.zeros(2)

    def state_dict(self):
        return dict(v=self.v.clone())

    def load_state_dict(self, sd):
        self.v.copy_(sd["v"])


class _StubModel:
    """train_step all-reduces a 'gradient' like FlatAdamW.all_reduce_grads does, so ranks must stay in lock-step."""

    def __init__(self):
        self.opt = _StubOpt()
        self.generator = torch.Generator().manual_seed(0)
        self.seen = []
        self.trainer = type("T", (), {"iteration": 0, "write_back": lambda s: None, "opt": self.opt})()

    def build_trainer(self, optimizer_cfg, lr_config=None, ema_cfg=None):
        self.ema_cfg = ema_cfg
        return {"diffusion": self.opt}

    def train_step(self, data, optimizer, running_status=None):
        ids = data["prompt_embed_kwargs"]["encoder_hidden_states"][:, 0, 0]
        self.seen += [float(v) for v in ids]
        g = ids.sum().reshape(1).clone()
        dist.all_reduce(g)
        optimizer["diffusion"].v += g / dist.get_world_size()
        return dict(log_vars=dict(loss=float(g)), num_samples=len(ids))

    def state_dict(self, trainable_only=True):
        return {"diffusion.denoising.w": self.opt.v.clone()}


class _IdDataset(torch.utils.data.Dataset):
    def __len__(self):
        return 16

    def __getitem__(self, i):
        return dict(prompt_embed_kwargs=dict(encoder_hidden_states=torch.full((2, 2), float(i))), latents=torch.zeros(16, 2, 2))


def _train_worker(rank, world, port, work_dir, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lakonlab.apis import train_model
        from lakonlab.utils import Config
        cfg = Config(dict(
            data=dict(workers_per_gpu=0, train_dataloader=dict(samples_per_gpu=2)), seed=3, module_wrapper="ddp",
            optimizer=dict(diffusion=dict(type="AdamW8bit", lr=1e-4)), lr_config=dict(policy="fixed"),
            runner=dict(type="DynamicIterBasedRunnerMod", pass_training_status=True), work_dir=work_dir, total_iters=4,
            custom_hooks=[dict(type="ExponentialMovingAverageHookMod", module_keys=("diffusion_ema",), start_iter=2,
                               momentum_policy="karras", momentum_cfg=dict(gamma=7.0))],
            checkpoint_config=dict(interval=2, max_keep_ckpts=-1, out_dir=os.path.join(work_dir, "ck")), name="stub",
            log_config=dict(interval=1, hooks=[dict(type="TextLoggerHook")]), workflow=[("train", 2)],
            resume_from=os.path.join(work_dir, "ck", "stub", "latest.pth")))
        model = _StubModel()
        runner = train_model(model, [_IdDataset()], cfg, distributed=True)
        q.put((rank, runner.iter, sorted(model.seen), model.opt.v.tolist(), model.ema_cfg))
    finally:
        dist.destroy_process_group()


def test_train_model_world2(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, it0, seen0, v0, ema0), (r1, it1, seen1, v1, ema1) = res
    assert it0 == it1 == 4
    assert len(seen0) == len(seen1) == 8 and not set(seen0) & set(seen1)      # DistributedSampler: disjoint shards
    assert v0 == v1                                                           # the all-reduced update is identical
    assert ema0 == dict(start_iter=2, momentum_cfg=dict(gamma=7.0))           # EMA hook config reaches the trainer
    ck = tmp_path / "ck" / "stub"
    assert sorted(p.name for p in ck.iterdir()) == ["iter_2.pth", "iter_4.pth", "latest.pth"]   # written once, by rank 0


# ------------------------------------------------------------------------------------------------
# Sharding through the PRODUCT API: pipe(...) called with the same arguments on every rank splits the batch, denoises its
# shard and all-gathers the final latents (lakonlab/pipelines/arcflux_pipeline.py). The engine is replaced by a CPU stub
# whose "denoise" is a per-image function, so what is tested is the host protocol: who gets which images, what noise.
# ------------------------------------------------------------------------------------------------
class _StubTransformer:
    num_gaussians = 16
    device = torch.device("cpu")
    cfg = type("C", (), {"in_channels": 64})()

    def __init__(self):
        self.seen = []

    def set_lora_scale(self, s):
        pass

    def denoise(self, latents, prompt_embeds, pooled, grid, **kw):
        self.seen.append(latents.shape[0])
        return latents * 2.0 + prompt_embeds.float().mean(dim=(1, 2))[:, None, None] + pooled.float().sum(1)[:, None, None]


def _pipe_call(batch, seed, with_latents):
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline
    g = torch.Generator().manual_seed(seed)
    txt, pooled = torch.randn(batch, 4, 8, generator=g), torch.randn(batch, 8, generator=g)
    kw = dict(latents=torch.randn(batch, 16, 64, generator=g)) if with_latents else dict(
        generator=torch.Generator().manual_seed(seed + 1))
    tr = _StubTransformer()
    pipe = ArcFluxPipeline(transformer=tr)
    out = pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, height=64, width=64, num_inference_steps=2,
               timestep_ratio=1.0, output_type="latent", **kw).images
    return out, tr.seen


def _pipe_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        for batch in (5, 2, 1):
            for with_latents in (True, False):
                out, seen = _pipe_call(batch, 11, with_latents)
                res[(batch, with_latents)] = (out, seen)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_pipeline_call_shards_the_batch_and_is_rank_count_invariant():
    single = {(b, w): _pipe_call(b, 11, w) for b in (5, 2, 1) for w in (True, False)}
    for (b, w), (out, seen) in single.items():
        assert seen == [b] and out.shape == (b, 16, 64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipe_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for key, (want, _) in single.items():
        b = key[0]
        for rank in (0, 1):
            out, seen = res[rank][key]
            assert torch.equal(out, want), (key, rank)          # every rank returns the FULL batch, same as one process
            lo, hi = shard_bounds(b, rank, 2)
            assert seen == ([hi - lo] if hi > lo else [])       # ... having denoised only its own images


def _rng_worker(rank, world, port, path, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lakonlab.runner.checkpoint import get_checkpoint, restore_rng_state, write_checkpoint_to_file, load_checkpoint
        model = type("M", (), {})()
        model.generator = torch.Generator().manual_seed(100 + rank)       # train.py --diff_seed: seed + rank
        model.state_dict = lambda trainable_only=True: {}
        torch.randn(3, generator=model.generator)                         # advance the streams
        ck = get_checkpoint(model)                                        # collective
        expect = torch.randn(4, generator=model.generator)                # what the uninterrupted run draws next
        if rank == 0:
            write_checkpoint_to_file(ck, path, create_symlink=False)
        dist.barrier()
        fresh = torch.Generator().manual_seed(0)
        how = restore_rng_state(fresh, load_checkpoint(path)["rng_state"], rank, 7)
        q.put((rank, how, torch.equal(torch.randn(4, generator=fresh), expect), expect.tolist()))
    finally:
        dist.destroy_process_group()


def test_checkpoint_keeps_every_ranks_rng_stream(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    path = str(tmp_path / "ck.pth")
    procs = [ctx.Process(target=_rng_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(how == "restored" and same for _, how, same, _ in res)     # each rank continues ITS OWN stream
    assert res[0][3] != res[1][3]                                         # and the streams differ between ranks
    # a rank the checkpoint does not know is re-seeded (distinct from every saved stream), an old bare-tensor state is rank 0's
    from lakonlab.runner.checkpoint import load_checkpoint, restore_rng_state
    saved = load_checkpoint(path)["rng_state"]
    g2, g3 = torch.Generator(), torch.Generator()
    assert restore_rng_state(g2, saved, 2, 7) == "reseeded" and restore_rng_state(g3, saved, 3, 7) == "reseeded"
    assert not torch.equal(torch.randn(4, generator=g2), torch.randn(4, generator=g3))
    g0 = torch.Generator()
    assert restore_rng_state(g0, saved[0], 0) == "restored"
