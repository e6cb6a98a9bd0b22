"""CPU: the block-wise 8-bit AdamW restatement (oracle/adamw8bit_oracle.py — bitsandbytes' AdamW8bit, the optimizer the
reference's configs name) against the values its code book is known to contain, against fp32 AdamW on a toy problem, and the
host-side arena layout of the 8-bit mode."""
import numpy as np
import pytest
import torch

from oracle import adamw8bit_oracle as O


def test_dynamic_map_known_values_and_shape():
    s, u = O.dynamic_map(True), O.dynamic_map(False)
    for m in (s, u):
        assert m.shape == (256,) and m.dtype == np.float32 and (np.diff(m) > 0).all()
        assert m[-1] == 1.0 and (m == 0).sum() == 1
    # signed: 127 negative fractions, 0, 127 positive fractions, 1; decades 1e-6 .. 1 with 1, 2, 4 .. 64 fraction steps
    assert s[127] == 0.0 and (s[:127] == -s[128:255][::-1]).all()
    assert s[-2] == pytest.approx(0.99296875, rel=1e-6)             # mid-point of the last of 64 steps of 0.9 / 64
    assert s[128] == pytest.approx(5.5e-7, rel=1e-6)                # mid-point of [0.1, 1] x 1e-6
    assert ((s > 0.1) & (s < 1.0)).sum() == 64
    # unsigned: twice the fraction steps per decade
    assert u[0] == 0.0 and u[1] == pytest.approx(3.25e-7, rel=1e-6) and ((u > 0.1) & (u < 1.0)).sum() == 128
    assert u[-2] == pytest.approx(1.0 - 0.9 / 128 / 2, rel=1e-6)


def test_product_code_books_equal_the_oracle():
    from arcflow_b200.optim import create_dynamic_map
    for signed in (True, False):
        np.testing.assert_allclose(create_dynamic_map(signed).numpy(), O.dynamic_map(signed), rtol=3e-7, atol=0)


def test_nearest_code_and_blockwise_round_trip():
    q = O.dynamic_map(True)
    x = np.array([-1.0, -0.5, 0.0, 1e-9, 0.3, 0.99, 1.0], np.float32)
    idx = O.nearest_code(q, x)
    brute = np.abs(q[None, :] - x[:, None]).argmin(axis=1)
    assert (np.abs(q[idx] - x) <= np.abs(q[brute] - x) + 1e-12).all()
    assert q[idx[2]] == 0.0 and q[idx[-1]] == 1.0 and idx[0] == 0
    rng = np.random.default_rng(0)
    v = (rng.standard_normal(4096) * 3e-3).astype(np.float32)
    codes, absmax = O.quantize_blockwise(v, q)
    back = O.dequantize_blockwise(codes, absmax, q)
    assert absmax.shape == (16,) and codes.dtype == np.uint8
    # dynamic quantisation: in the top decade (|x| > 0.1 absmax) the codes are 0.9 / 64 apart -> error <= half a step of the
    # block's absmax; everywhere the error stays below that bound (finer steps in the lower decades)
    am = np.repeat(absmax, 256)
    assert (np.abs(back - v) <= (0.9 / 128) * am * 1.001).all()
    top = np.abs(v) > 0.1 * am
    assert top.sum() > 1000 and (np.abs(back - v)[top] <= 0.0704 * np.abs(v)[top]).all()


def test_8bit_adamw_tracks_fp32_adamw_on_a_regression():
    rng = np.random.default_rng(1)
    n, rows = 512, 2048                                             # over-determined: one optimum
    A = rng.standard_normal((rows, n)).astype(np.float32) / 16
    target = (A @ rng.standard_normal(n)).astype(np.float32)
    q1, q2 = O.dynamic_map(True), O.dynamic_map(False)
    p8 = np.zeros(n, np.float32)
    p32 = p8.copy()
    c1 = np.zeros(n, np.uint8); c2 = np.zeros(n, np.uint8)
    a1 = np.zeros(n // 256, np.float32); a2 = a1.copy()
    m = np.zeros(n, np.float32); v = m.copy()
    loss = lambda p: float(((A @ p - target) ** 2).mean())
    l0 = loss(p8)
    for t in range(1, 601):
        g8 = (2 * A.T @ (A @ p8 - target) / rows * 16).astype(np.float32)
        g32 = (2 * A.T @ (A @ p32 - target) / rows * 16).astype(np.float32)
        p8, c1, c2, a1, a2 = O.adamw8bit_step(p8, g8, c1, c2, a1, a2, q1, q2, t, 1e-2)
        p32, m, v = O.adamw32_step(p32, g32, m, v, t, 1e-2)
    assert loss(p32) < 1e-3 * l0 and loss(p8) < 1e-3 * l0          # both reach the optimum; quantisation noise only
    assert np.linalg.norm(p8 - p32) / np.linalg.norm(p32) < 0.05


def test_first_step_from_zero_state_is_a_sign_step():
    """t = 1 from zero moments: s1 = (1 - b1) g, s2 = (1 - b2) g^2 -> update = -lr * sign(g) (up to eps), whatever the codes."""
    q1, q2 = O.dynamic_map(True), O.dynamic_map(False)
    g = np.linspace(-2, 2, 256).astype(np.float32)
    g[g == 0] = 0.5
    p, c1, c2, a1, a2 = O.adamw8bit_step(np.zeros(256, np.float32), g, np.zeros(256, np.uint8), np.zeros(256, np.uint8),
                                         np.zeros(1, np.float32), np.zeros(1, np.float32), q1, q2, 1, 1e-3)
    np.testing.assert_allclose(p, -1e-3 * np.sign(g), rtol=1e-4)
    assert a1[0] == pytest.approx(0.1 * 2.0, rel=1e-6) and a2[0] == pytest.approx(0.05 * 4.0, rel=1e-5)
    assert (np.sign(q1[c1]) == np.sign(g)).all()                    # exp_avg never loses its sign in the code book


def test_flat_arena_layout_of_the_8bit_mode():
    from arcflow_b200.optim import FlatAdamW, MIN_8BIT_SIZE, QBLOCK
    shapes = {"blocks.0.lora_A.weight": (16, 300), "blocks.0.lora_B.weight": (300, 16), "proj_out_means.weight": (64, 100),
              "proj_out_means.bias": (64,), "proj_out_loggamma.weight": (60, 100), "proj_out_loggamma.bias": (60,),
              "norm_out.linear.bias": (200,)}
    o = FlatAdamW(shapes, "cpu", state_bits=8)
    numel = lambda n: int(np.prod(shapes[n]))
    big = [n for n in shapes if numel(n) >= MIN_8BIT_SIZE]
    small = [n for n in shapes if numel(n) < MIN_8BIT_SIZE]
    assert o.n8 % QBLOCK == 0 and o.n8 > 0 and o.n > o.n8
    for n in big:        # 8-bit tensors: block-aligned slots inside [0, n8), never sharing a block
        off = o.views[n][0]
        assert off % QBLOCK == 0 and off + numel(n) <= o.n8
    for n in small:      # fp32-state tensors behind the seam
        assert o.views[n][0] >= o.n8 and o.views[n][0] % 4 == 0
    spans = sorted((o.views[n][0], o.views[n][0] + numel(n)) for n in shapes)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))     # no overlap
    lo = [n for n in shapes if "proj_out_loggamma" in n]
    inside = [n for n in shapes if o.lo[0] <= o.views[n][0] < o.lo[1]]
    assert sorted(inside) == sorted(lo) and o.lo[0] % QBLOCK == 0  # the lr-multiplier range holds exactly those tensors
    assert o.state1.numel() == o.n8 and o.absmax1.numel() == o.n8 // QBLOCK and o.exp_avg.numel() == o.n - o.n8
    # the fp32 mode keeps its round-1 layout (checkpoints stay loadable)
    o32 = FlatAdamW(shapes, "cpu")
    assert o32.n8 == 0 and o32.lo[0] == 0 and o32.exp_avg.numel() == o32.n
    with pytest.raises(Exception, match="8-bit|bit"):
        o32.load_state_dict(o.state_dict())
