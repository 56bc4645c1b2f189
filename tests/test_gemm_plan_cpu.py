"""Tile schedule of the CTA-pair GEMM for the model's shapes (ViT-B/16@384, batch 16 → M = 9232) — host logic only.

The kernel deals work items round-robin to 74 CTA pairs (148 SMs); these tests pin the decisions DESIGN.md §4
describes: 256-wide tiles where the waves are full, mixed 256+128 tiles or 192-wide tiles for N = 768, split-K only
for the weight-gradient (accumulating) epilogue."""
import pytest

import chest_x_ray_vit_b200 as pkg

ops = pkg.ops
M, D, F = 9232, 768, 3072


def test_wide_forward_gemms_use_256_tiles():
    fc1 = ops.gemm_plan(M, F, D, epilogue=ops.EPI_BIAS_GELUG_BF16)
    assert fc1 == {"tile_n": 256, "split_k": 1, "n_half": 0, "work_items": 37 * 12}      # 444 = 74 pairs × 6 tiles
    # N = 2304: 333 tiles of 256 would be 4.5 waves (5 tiles on the busiest pairs); one 256-wide tile per band is
    # replaced by two 128-wide ones → 296 full + 74 half items = 4 full tiles and 1 half tile on every pair
    qkv = ops.gemm_plan(M, 3 * D, D, epilogue=ops.EPI_BIAS_BF16)
    assert qkv == {"tile_n": 256, "split_k": 1, "n_half": 2, "work_items": 37 * 10}


def test_n768_balances_the_static_schedule():
    # K-major B (forward): 192-wide tiles → 37 bands × 4 = 148 tiles = exactly two per pair
    fc2 = ops.gemm_plan(M, D, F, epilogue=ops.EPI_BIAS_RESID_F32)
    assert fc2["tile_n"] == 192 and fc2["n_half"] == 0 and fc2["work_items"] == 148
    # MN-major B (dgrad): 192 is not a whole number of swizzle atoms per CTA → 2 tiles of 256 + 2 of 128 per band
    for K in (F, 3 * D, D):
        dg = ops.gemm_plan(M, D, K, epilogue=ops.EPI_STORE_BF16, b_mn_major=True)
        assert dg == {"tile_n": 256, "split_k": 1, "n_half": 2, "work_items": 148}, (K, dg)
    # a forced tile width switches the mixed schedule off
    assert ops.gemm_plan(M, D, F, epilogue=ops.EPI_STORE_BF16, b_mn_major=True, tile_n=256)["n_half"] == 0


def test_weight_gradients_split_k_to_fill_the_pairs():
    for (m, n) in ((D, F), (F, D)):
        wg = ops.gemm_plan(m, n, M, epilogue=ops.EPI_ACCUM_F32, a_mn_major=True, b_mn_major=True)
        assert wg["tile_n"] == 256 and wg["n_half"] == 0 and wg["split_k"] == 2 and wg["work_items"] == 72
    small = ops.gemm_plan(D, D, M, epilogue=ops.EPI_ACCUM_F32, a_mn_major=True, b_mn_major=True)
    assert small["split_k"] >= 4 and small["work_items"] <= 74
    # only the accumulating epilogue may split K
    assert ops.gemm_plan(D, D, M, epilogue=ops.EPI_STORE_BF16)["split_k"] == 1


def test_fewer_sms_rebalance_and_errors():
    # SMs left to a concurrent collective: the schedule is recomputed for the remaining pairs
    capped = ops.gemm_plan(M, 3 * D, D, epilogue=ops.EPI_BIAS_BF16, max_ctas=128)
    assert capped["tile_n"] in (128, 192, 256) and capped["work_items"] >= 37 * 9
    with pytest.raises(RuntimeError):
        ops.gemm_plan(M, 100, D, epilogue=ops.EPI_STORE_BF16)           # N must be a multiple of 128

