"""Host-side data-layout logic of the streamed decode path (pure torch, runs without a GPU)."""
import torch

from video2music_b200.engine import pack_fragments, swizzled_er_copies


def test_pack_fragments_matches_mma_a_fragment_layout():
    """Lane l = 4g + q of tile (n16, k16) holds rows (g, g+8) x columns (2q, 2q+1, 2q+8, 2q+9) in the register order
    of mma.m16n8k16's A operand: [k-half][row-half][pair]."""
    N, K = 40, 64                                               # N is padded to 48 with zero rows
    w = torch.arange(N * K, dtype=torch.float32).view(N, K)
    wb = w.to(torch.bfloat16)
    P = pack_fragments(wb)
    assert P.shape == (3, 4, 32, 8) and P.dtype == torch.bfloat16
    ref = torch.cat([wb, wb.new_zeros(8, K)], 0)
    for n16 in range(3):
        for k16 in range(4):
            for lane in range(32):
                g, q = lane >> 2, lane & 3
                r0, c0 = n16 * 16 + g, k16 * 16 + 2 * q
                exp = torch.stack([ref[r0, c0], ref[r0, c0 + 1], ref[r0 + 8, c0], ref[r0 + 8, c0 + 1],
                                   ref[r0, c0 + 8], ref[r0, c0 + 9], ref[r0 + 8, c0 + 8], ref[r0 + 8, c0 + 9]])
                assert torch.equal(P[n16, k16, lane], exp)


def test_swizzled_er_copies():
    """Copy s stores row r with chunk c at position c ^ ((r - s) & 7): a slice starting at any row `start`, taken from
    copy start & 7, has chunk c of its j-th row at position c ^ (j & 7)."""
    er = torch.arange(300 * 64, dtype=torch.float32).view(300, 64).to(torch.bfloat16)
    sw = swizzled_er_copies(er)
    assert sw.shape == (8, 300, 64)
    for start in (0, 1, 7, 8, 13, 150, 299 - 20):
        sl = sw[start & 7, start:start + 20].view(20, 8, 8)
        for j in range(20):
            for c in range(8):
                assert torch.equal(sl[j, c ^ (j & 7)], er[start + j].view(8, 8)[c])
