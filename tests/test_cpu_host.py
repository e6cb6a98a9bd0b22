"""CPU tier: the C-ABI library loads and exports every declared symbol, and the host-side logic
(schedule, RoPE tables, weight packing, argument validation) agrees with the oracle. No compute calls."""
import re
from pathlib import Path

import pytest
import torch

from oracle import arcflow_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(lib):
    from arcflow_b200 import _lib
    header = (ROOT / "include" / "arcflow_b200.h").read_text()
    declared = set(re.findall(r"\b(afb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libarcflow_b200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"
    assert lib.afb_abi_version() == _lib.AFB_ABI_VERSION
    assert lib.afb_launch_count() == 0


def test_error_path_without_gpu(lib):
    """A null descriptor must come back as an error code + message, never a crash."""
    from arcflow_b200 import _lib
    rc = lib.afb_gemm(None, None)
    assert rc == -1
    assert b"null" in lib.afb_last_error()
    with pytest.raises(_lib.AfbError):
        _lib.check(rc, "afb_gemm")


def test_ops_refuse_cpu_tensors(lib):
    from arcflow_b200 import ops, AfbError
    a = torch.zeros(1, 128, 64, dtype=torch.bfloat16)
    with pytest.raises(AfbError, match="CUDA"):
        ops.gemm(a, torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(1, 128, 8, dtype=torch.bfloat16))
    with pytest.raises(AfbError, match="CUDA"):
        ops.attention(a, a, a)


@pytest.mark.parametrize("nfe,ratio", [(1, 1.0), (2, 1.0), (4, 1.0), (8, 1.0), (4, 0.5), (2, 0.5)])
def test_schedule_matches_oracle(nfe, ratio):
    from arcflow_b200 import schedule
    raw, sub, tot = O.retrieve_raw_timesteps(nfe, 128, ratio)
    assert schedule.retrieve_raw_timesteps(nfe, 128, ratio) == (raw, sub, tot)
    ts = O.scheduler_timesteps(raw, 3.2)
    sig = schedule.denoise_sigmas(nfe, 128, ratio, 3.2)
    idx = 0
    for i in range(nfe):
        assert sig[i] == float((ts[idx] / 1000.0).item())
        idx += sub[i]
    assert sig[-1] == 0.0 and len(sig) == nfe + 1


def test_flux_time_quirk_values():
    # what the time embedder sees after `timestep.to(bf16) * 1000`
    from arcflow_b200.schedule import denoise_sigmas, flux_time_inputs
    seen = [flux_time_inputs(s, 3.5)[0] for s in denoise_sigmas(8)[:-1]]
    assert seen == [1000, 956, 908, 844, 760, 656, 516, 314]
    assert flux_time_inputs(1.0, 3.5)[1] == 3504


@pytest.mark.parametrize("txt,gh,gw", [(512, 16, 16), (7, 3, 5), (64, 64, 64)])
def test_rope_tables_match_oracle(txt, gh, gw):
    from arcflow_b200.rope import flux_rope_tables
    cos, sin = flux_rope_tables(txt, gh, gw, round_bf16=False)
    oc, os_ = O.flux_rope(txt, gh, gw)
    assert torch.equal(cos, oc) and torch.equal(sin, os_)
    cb, sb = flux_rope_tables(txt, gh, gw, round_bf16=True)
    assert torch.equal(cb, oc.bfloat16().float()) and torch.equal(sb, os_.bfloat16().float())


def test_packing_layouts_are_algebraically_equivalent():
    """[x | x A^T] @ [W | B]^T == x W^T + (x A^T) B^T, fused QKV == the three projections, the modulation
    concat keeps each block's chunk order, and the fused head keeps (means | logits | loggamma | pad)."""
    import ctypes as C
    from arcflow_b200.config import flux_tiny
    from arcflow_b200.model import PackedFluxWeights
    from arcflow_b200.synthetic import make_flux_state_dict
    cfg = flux_tiny(1, 1, 2)
    sd = make_flux_state_dict(cfg, seed=3)
    pk = PackedFluxWeights(dict(sd), cfg, device="cpu")
    by_ptr = {t.data_ptr(): t for t in pk.keep}
    D, M, r = cfg.inner_dim, cfg.mlp_dim, cfg.lora_rank
    x = torch.randn(5, D).bfloat16()

    k = pk.dbl[0]
    up = by_ptr[k.img_up_w].float()
    la = by_ptr[k.img_up_la].float()
    assert up.shape == (M, D + r) and la.shape == (r, D)
    ref = O._lin(sd, "transformer_blocks.0.ff.net.0.proj", x.float(), torch.float32)
    t = x.float() @ la.t()
    got = torch.cat([x.float(), t], 1) @ up.t() + by_ptr[k.img_up_b].float()
    assert torch.allclose(got, ref, atol=1e-4)

    qkv = by_ptr[k.txt_qkv_w]
    for j, n in enumerate(("add_q_proj", "add_k_proj", "add_v_proj")):
        assert torch.equal(qkv[j * D:(j + 1) * D], sd[f"transformer_blocks.0.attn.{n}.weight"])
    s = pk.sgl[0]
    assert by_ptr[s.out_w].shape == (D, D + M + r) and by_ptr[s.out_la].shape == (r, D + M)

    mod_w = by_ptr[pk.struct.mod_w]
    assert mod_w.shape[0] == pk.struct.mod_total == 6 * D * 2 + 3 * D + 2 * D
    assert torch.equal(mod_w[k.txt_mod_off:k.txt_mod_off + 6 * D], sd["transformer_blocks.0.norm1_context.linear.weight"])
    assert torch.equal(mod_w[s.mod_off:s.mod_off + 3 * D], sd["single_transformer_blocks.0.norm.linear.weight"])
    assert torch.equal(mod_w[pk.struct.norm_out_mod_off:], sd["norm_out.linear.weight"])

    head = by_ptr[pk.struct.head_w]
    assert head.shape == (1152, D) and pk.struct.head_n == 1152
    assert torch.equal(head[:1024], sd["proj_out_means.weight"])
    assert torch.equal(head[1024:1088], sd["proj_out_logweights.weight"])
    assert torch.equal(head[1088:1148], sd["proj_out_loggamma.weight"])
    assert head[1148:].abs().sum() == 0


def test_packing_reports_missing_keys():
    from arcflow_b200 import AfbError
    from arcflow_b200.config import flux_tiny
    from arcflow_b200.model import PackedFluxWeights
    from arcflow_b200.synthetic import make_flux_state_dict
    cfg = flux_tiny(1, 1, 2)
    sd = make_flux_state_dict(cfg, seed=3)
    del sd["proj_out_loggamma.weight"]
    with pytest.raises(AfbError, match="proj_out_loggamma.weight"):
        PackedFluxWeights(sd, cfg, device="cpu")


def test_synthetic_loggamma_bias_is_reference_init():
    # lakonlab/models/architecture/arcflow/arcflux.py:115-132; values in SURVEY.md Appendix D
    from arcflow_b200.config import flux_tiny
    from arcflow_b200.synthetic import make_flux_state_dict
    sd = make_flux_state_dict(flux_tiny(1, 1, 2), dtype=torch.float32)
    b = sd["proj_out_loggamma.bias"].reshape(15, 4)
    assert torch.allclose(b[:, 0], b[:, 3])
    assert b[0, 0].item() == pytest.approx(-1.6094, abs=1e-4)
    assert b[14, 0].item() == pytest.approx(1.3863, abs=1e-4)


def test_oracle_forward_shapes_and_bf16_agreement():
    """The oracle itself: fp32 vs the reference's bf16 numerics stay within bf16 noise on a tiny model."""
    from arcflow_b200.config import flux_tiny
    from arcflow_b200.synthetic import make_flux_state_dict, make_flux_inputs
    cfg = flux_tiny(1, 1, 2)
    sd = make_flux_state_dict(cfg, seed=5)
    x, txt, pooled = make_flux_inputs(cfg, 1, 32, 32, txt_len=16)
    args = (x.bfloat16(), txt, pooled, torch.tensor([0.76]), torch.tensor([3.5]), (2, 2))
    a = O.flux_forward(sd, cfg, *args, dtype=torch.float32)
    b = O.flux_forward(sd, cfg, *args, dtype=torch.bfloat16)
    assert a["means"].shape == (1, 4, 16, 64) and a["logweights"].shape == (1, 4, 16, 4)
    assert a["loggammas"].shape == (1, 4, 15, 4)
    assert torch.allclose(a["logweights"].exp().sum(-2), torch.ones(1, 4, 4), atol=1e-5)
    err = ((a["means"] - b["means"].float()).norm() / a["means"].norm()).item()
    assert err < 5e-2


@pytest.mark.parametrize("txt,gh,gw", [(512, 64, 64), (33, 3, 5), (7, 8, 2)])
def test_qwen_rope_tables_match_oracle(txt, gh, gw):
    from arcflow_b200.qwen import qwen_rope_tables
    cos, sin = qwen_rope_tables(txt, gh, gw)
    vid, tx = O.QwenEmbedRope()(1, gh, gw, txt)
    fc = torch.cat([tx, vid], 0)            # text first in the joint sequence
    assert cos.shape == (txt + gh * gw, 128)
    for tab, ref in ((cos, fc.real), (sin, fc.imag)):
        assert torch.allclose(tab[:, 0::2], ref, atol=2e-7) and torch.equal(tab[:, 0::2], tab[:, 1::2])


def test_qwen_packing_and_lora_targets():
    from arcflow_b200.qwen import PackedQwenWeights, make_qwen_state_dict, qwen_lora_targets, qwen_tiny, qwen_time_input
    cfg = qwen_tiny(3, 2)
    sd = make_qwen_state_dict(cfg, seed=9)
    # configs/qwen/arcqwen_2nfe_k16.py:53-55: txt_mlp of the last block carries no LoRA
    assert "transformer_blocks.2.txt_mlp.net.2.lora_A.weight" not in sd
    assert "transformer_blocks.1.txt_mlp.net.2.lora_A.weight" in sd
    assert len(qwen_lora_targets(cfg)) == 2 + 3 * 2 + 2 * 2
    pk = PackedQwenWeights(dict(sd), cfg, device="cpu")
    by_ptr = {t.data_ptr(): t for t in pk.keep}
    D, M, r = cfg.inner_dim, cfg.mlp_dim, cfg.lora_rank
    assert by_ptr[pk.dbl[0].txt_up_w].shape == (M, D + r)
    assert by_ptr[pk.dbl[2].txt_up_w].shape == (M, D) and not pk.dbl[2].txt_up_la
    assert pk.struct.mod_total == 3 * 12 * D + 2 * D
    assert qwen_time_input(0.7619047761) == pytest.approx(761.71875)      # bf16(sigma) * 1000, SURVEY App. D


def test_bench_stdout_carries_only_the_result_line():
    """bench.py's contract is ONE JSON line on stdout; library chatter (NCCL's version banner goes through C stdio) must land
    on stderr even when it is flushed at exit."""
    import subprocess
    import sys as _sys
    import textwrap
    code = textwrap.dedent('''
        import ctypes, sys
        sys.argv = ["bench.py"]
        import bench
        bench._claim_stdout()
        ctypes.CDLL(None).puts(b"NCCL version 0.0.0 (C stdio)")
        print("python-level chatter")
        bench._emit({"metric": "m", "value": 1.0})
    ''')
    r = subprocess.run([_sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": "m", "value": 1.0}\n'
    assert "NCCL version 0.0.0" in r.stderr and "python-level chatter" in r.stderr
