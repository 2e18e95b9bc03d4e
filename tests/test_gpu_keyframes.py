"""GPU parity: K1 (nested masks + Interp), interpolate_from_indices, corruption builders -- CUDA path through
the C ABI vs the oracle on the same seeded inputs, vs the committed golden vectors, plus size-independent
properties at BASELINE.json's full size (1M trajectories).  Bars: bit-exact (masks, idx, fp32 interpolation)."""
import numpy as np
import pytest
import torch

from oracle import keyframes_np as okf
from oracle import sampling_np as osp

pytestmark = pytest.mark.gpu
SCHED = {0: "doubling", 1: "linear", 2: "geom"}


@pytest.fixture(scope="module")
def kf():
    from interpolated_diffusion_b200.corruptions import keyframes
    return keyframes


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def eq(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b), f"mismatch: {np.sum(a != b)} of {a.size}"


def test_nested_masks_golden(kf, golden):
    g = golden("keyframes")
    for tag in "abcdetfg":
        B, T, K, S, sc = [int(v) for v in g[f"nest_{tag}_cfg"]]
        masks, idxs = kf.build_nested_masks_batch(B, T, K, S, k_schedule=SCHED[sc], scores=dev(g[f"nest_{tag}_scores"]))
        assert masks.dtype == torch.bool and idxs[0].dtype == torch.int64
        eq(masks, g[f"nest_{tag}_masks"])
        for s in range(S + 1):
            eq(idxs[s], g[f"nest_{tag}_idx{s}"])


def test_fixed_k_and_uniform_golden(kf, golden):
    g = golden("keyframes")
    for tag in "abcde":
        B, T, K, ee = [int(v) for v in g[f"fixk_{tag}_cfg"]]
        idx, mask = kf.sample_fixed_k_indices_batch(B, T, K, ensure_endpoints=bool(ee), scores=dev(g[f"fixk_{tag}_scores"]))
        eq(idx, g[f"fixk_{tag}_idx"])
        eq(mask, g[f"fixk_{tag}_mask"])
    for row in g["uniform"]:
        T, K = int(row[0]), int(row[1])
        idx, mask = kf.sample_fixed_k_indices_uniform_batch(3, T, K, device="cuda")
        eq(idx[2], row[2:2 + K])
    for tag in "ab":
        B, T, K, jit = g[f"unij_{tag}_cfg"]
        idx, mask = kf.sample_fixed_k_indices_uniform_batch(int(B), int(T), int(K), device="cuda", jitter=float(jit),
                                                             u=dev(g[f"unij_{tag}_u"]))
        eq(idx, g[f"unij_{tag}_idx"])
        eq(mask, g[f"unij_{tag}_mask"])


def test_from_logits_and_level_logits_golden(kf, golden):
    g = golden("keyframes")
    masks, idxs = kf.build_nested_masks_from_logits(dev(g["logit_logits"]), 4, 2)
    eq(masks, g["logit_masks"])
    for s in range(3):
        eq(idxs[s], g[f"logit_idx{s}"])
    masks, idxs = kf.build_nested_masks_from_level_logits(dev(g["lvl_logits"]), 4, 2)
    eq(masks, g["lvl_masks"])
    for s in range(3):
        eq(idxs[s], g[f"lvl_idx{s}"])


def test_from_base_invariants(kf):
    # the reference's sequential randperm cannot be replayed in parallel: check the contract
    B, T, K, S = 512, 64, 8, 3
    idx_base, _ = kf.sample_fixed_k_indices_uniform_batch(B, T, K, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    masks, idxs = kf.build_nested_masks_from_base(idx_base, T, S, generator=gen)
    K_list = kf._compute_k_schedule(T, K, S)
    eq(idxs[S], idx_base)
    m = masks.cpu().numpy()
    for s in range(S + 1):
        assert (m[:, s].sum(1) == K_list[s]).all()
        assert idxs[s].shape == (B, K_list[s])
        assert (torch.diff(idxs[s], dim=1) > 0).all()
    for s in range(1, S + 1):
        assert (m[:, s] <= m[:, s - 1]).all()
    # injected scores: oracle equivalence (base first, then stable ascending rank of the scores)
    sc = torch.rand((B, T), generator=gen, device="cuda")
    masks2, _ = kf.build_nested_masks_from_base(idx_base, T, S, scores=sc)
    scn = sc.cpu().numpy().copy()
    np.put_along_axis(scn, idx_base.cpu().numpy(), -1.0, axis=1)
    rank = okf.stable_rank(scn)
    for s in range(S + 1):
        eq(masks2[:, s], rank < K_list[s])


def test_interpolate_from_indices_golden(kf, golden):
    g = golden("keyframes")
    for tag in "abcdefg":
        B, T, K, D, vel = [int(v) for v in g[f"interp_{tag}_cfg"]]
        y = kf.interpolate_from_indices(dev(g[f"interp_{tag}_idx"]), dev(g[f"interp_{tag}_vals"]), T, bool(vel))
        eq(y, g[f"interp_{tag}_y"])
    eq(kf.interpolate_from_indices(dev(g["interp_x_idx"]), dev(g["interp_x_vals"]), 16), g["interp_x_y"])
    ka = kf.interpolate_from_indices(torch.tensor([[0, 3, 6, 7]]).cuda(),
                                     torch.tensor([[[0.0], [3.0], [12.0], [7.0]]]).cuda(), 8)
    eq(ka[0, :, 0], np.array([0, 1, 2, 3, 6, 9, 12, 7], np.float32))
    # legacy mask API: <= 1 ulp from the loop implementation, exact against the oracle's idx form
    y = kf.interpolate_from_mask(dev(g["imask_x"]), dev(g["imask_m"]), False)
    eq(y, okf.interpolate_from_mask(g["imask_x"], g["imask_m"], False))
    np.testing.assert_allclose(y.cpu().numpy(), g["imask_y"], rtol=0, atol=2.4e-7)
    # reference tests/test_corruption.py:9-23
    x = torch.tensor([[0.0], [2.0], [4.0], [6.0], [8.0]]).cuda()
    m = torch.tensor([1, 0, 0, 0, 1], dtype=torch.bool).cuda()
    assert torch.allclose(kf.interpolate_keyframes(x, m), x)


@pytest.mark.parametrize("T,K,S,D,vel", [(64, 8, 3, 4, False), (64, 8, 3, 4, True), (64, 8, 3, 2, False),
                                          (256, 32, 4, 4, True), (256, 32, 4, 2, False), (128, 8, 3, 4, True),
                                          (32, 4, 2, 4, False), (16, 3, 3, 2, False), (100, 7, 3, 4, True),
                                          (33, 5, 2, 2, False)])
def test_fused_masks_interp_vs_oracle(kf, T, K, S, D, vel):
    B = 777
    gen = torch.Generator().manual_seed(T * 1000 + K)
    scores = torch.rand((B, T - 2), generator=gen)
    scores[5] = torch.floor(scores[5] * 4) / 4          # forced ties
    scores[6, 3] = scores[6, 17 % (T - 2)]
    x0 = torch.rand((B, T, D), generator=gen) * 3 - 1
    K_list = okf.compute_k_schedule(T, K, S)
    flags = kf.F_RECOMPUTE_VELOCITY if vel else 0
    masks, idxs, xl = kf.nested_masks_interp(scores.cuda(), T, K_list, x0=x0.cuda(), levels_out=(1, S), flags=flags)
    m_ref, i_ref = okf.build_nested_masks_batch(scores.numpy(), T, K, S)
    eq(masks, m_ref)
    rows = np.arange(B)[:, None]
    for s in range(S + 1):
        eq(idxs[s], i_ref[s])
    for s in range(1, S + 1):
        y_ref = okf.interpolate_from_indices(i_ref[s], x0.numpy()[rows, i_ref[s]], T, vel)
        eq(xl[s - 1], y_ref)


def test_full_size_properties(kf):
    """BASELINE.json config 2 at full size: B = 2^20, T = 64, D = 4, S = 3 (size-independent properties)."""
    B, T, D, K, S = 1 << 20, 64, 4, 8, 3
    gen = torch.Generator(device="cuda").manual_seed(1234)
    scores = torch.rand((B, T - 2), generator=gen, device="cuda")
    x0 = torch.rand((B, T, D), generator=gen, device="cuda")
    K_list = kf._compute_k_schedule(T, K, S)
    masks, idxs, xl = kf.nested_masks_interp(scores, T, K_list, x0=x0, levels_out=(1, S))
    torch.cuda.synchronize()
    counts = masks.sum(dim=2)
    for s in range(S + 1):
        assert (counts[:, s] == K_list[s]).all()
        assert (idxs[s][:, 0] == 0).all() and (idxs[s][:, -1] == T - 1).all()
        assert (torch.diff(idxs[s], dim=1) > 0).all()
        assert masks[:, s].gather(1, idxs[s]).all()
    for s in range(1, S + 1):
        assert (masks[:, s] <= masks[:, s - 1]).all()                       # nested
        y = xl[s - 1]
        m = masks[:, s].unsqueeze(-1)
        assert torch.equal(torch.where(m, y, 0), torch.where(m, x0, 0))     # anchors exact
        # idempotence: interpolating the interpolant from the same anchors reproduces it bit for bit
        vals = y.gather(1, idxs[s].unsqueeze(-1).expand(-1, -1, D))
        assert torch.equal(kf.interpolate_from_indices(idxs[s], vals, T), y)
    # linearity of Interp in x0 on a linear ramp: exact reproduction (reference tests/test_corruption.py:9-14)
    ramp = torch.arange(T, device="cuda", dtype=torch.float32).view(1, T, 1).expand(4096, T, D).contiguous()
    _, _, xr = kf.nested_masks_interp(scores[:4096], T, K_list, x0=ramp, levels_out=(S, S))
    assert torch.allclose(xr[0], ramp, rtol=0, atol=4e-6)
    # a slice agrees bit-for-bit with the CPU oracle
    n = 2048
    m_ref, i_ref = okf.build_nested_masks_batch(scores[:n].cpu().numpy(), T, K, S)
    eq(masks[:n], m_ref)
    rows = np.arange(n)[:, None]
    for s in range(1, S + 1):
        eq(xl[s - 1][:n], okf.interpolate_from_indices(i_ref[s], x0[:n].cpu().numpy()[rows, i_ref[s]], T))


def test_edge_cases(kf):
    # empty batch, T == 2, K >= T, ragged T
    masks, idxs = kf.build_nested_masks_batch(0, 16, 3, 2, device="cuda")
    assert masks.shape == (0, 3, 16) and idxs[0].shape[0] == 0
    masks, idxs = kf.build_nested_masks_batch(4, 2, 2, 1, device="cuda")
    assert masks.all() and idxs[0].tolist() == [[0, 1]] * 4
    masks, idxs = kf.build_nested_masks_batch(4, 8, 64, 2, device="cuda")
    assert masks.all()
    y = kf.interpolate_from_indices(torch.zeros((0, 3), dtype=torch.long).cuda(), torch.zeros((0, 3, 2)).cuda(), 8)
    assert y.shape == (0, 8, 2)
    with pytest.raises(RuntimeError):
        kf.nested_masks_interp(torch.rand(2, 510).cuda(), 512, [8])


def test_corrupt_from_anchors_golden(golden):
    from interpolated_diffusion_b200.train import train_interp_levels as tr
    g = golden("sampling")
    src, idx = dev(g["co_src"]), dev(g["co_idx"])
    for tag in "abcd":
        jit, jprob, mode, vel, sigma, asig = g[f"co_{tag}_cfg"]
        draws = [dev(g[f"co_{tag}_draw{i}"]) for i in range(int(g[f"co_{tag}_n"][0]))]
        pos = [0]
        o_randn, o_rand, o_randint = torch.randn, torch.rand, torch.randint

        def replay(*a, **k):
            d = draws[pos[0]]; pos[0] += 1
            return d

        torch.randn = torch.rand = torch.randint = replay
        try:
            out = tr._corrupt_from_anchors(src, idx, 32, None, float(sigma), float(asig), int(jit), float(jprob),
                                           "dist" if mode == 0 else "const", True, bool(vel))
        finally:
            torch.randn, torch.rand, torch.randint = o_randn, o_rand, o_randint
        assert pos[0] == len(draws)
        eq(out, g[f"co_{tag}_out"])


def test_build_interp_batches_golden(golden):
    from interpolated_diffusion_b200.train import train_interp_levels as tr
    g = golden("sampling")
    idx_levels = [dev(g[f"ad_idx{s}"]) for s in range(4)]
    x0, masks, s_idx = dev(g["ad_x0"]), dev(g["ad_masks"]), dev(g["ad_s_idx"])
    xs, xp, ms, mp, *_ = tr.build_interp_adjacent_batch(x0, 8, 3, None, masks_levels=masks, idx_levels=idx_levels, s_idx=s_idx)
    eq(xs, g["ad_none_xs"]); eq(xp, g["ad_none_xp"]); eq(ms, g["ad_none_ms"]); eq(mp, g["ad_none_mp"])
    xs, ms, *_ = tr.build_interp_level_batch(x0, 8, 3, None, masks_levels=masks, idx_levels=idx_levels, s_idx=s_idx)
    eq(xs, g["lv_none_xs"]); eq(ms, g["lv_none_ms"])
    draws = [dev(g[f"ad_dist_draw{i}"]) for i in range(int(g["ad_dist_n"][0]))]
    pos = [0]
    o_randn = torch.randn

    def replay(*a, **k):
        d = draws[pos[0]]; pos[0] += 1
        return d

    torch.randn = replay
    try:
        xs, xp, *_ = tr.build_interp_adjacent_batch(x0, 8, 3, None, masks_levels=masks, idx_levels=idx_levels, s_idx=s_idx,
                                                     corrupt_mode="dist", corrupt_sigma_max=0.08, corrupt_sigma_min=0.012,
                                                     corrupt_sigma_pow=0.75, corrupt_anchor_frac=0.25)
    finally:
        torch.randn = o_randn
    assert pos[0] == len(draws)
    eq(xs, g["ad_dist_xs"]); eq(xp, g["ad_dist_xp"])
