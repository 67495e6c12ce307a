"""CPU test of the seam bookkeeping used by sharded.resolve_flats_band: outlet flags are OR-ed and
flat heights MAX-ed between the pieces of a flat that meet at a band seam, until nothing changes.
The CUDA state object is replaced by plain CPU tensors (same attribute names)."""
import torch

from richdem_b200.sharded import CudaFlatsBand


def fake_band(h, w, gt, gb):
    b = object.__new__(CudaFlatsBand)
    b.h, b.w, b.gt, b.gb = h, w, gt, gb
    b.ft = torch.zeros((h, w), dtype=torch.uint8)
    b.root = torch.arange(h * w, dtype=torch.int32).reshape(h, w).clone()
    b.rootflag = torch.zeros(h * w, dtype=torch.uint8)
    b.height = torch.zeros(h * w, dtype=torch.int32)
    return b


def test_flags_cross_a_seam_both_ways():
    w = 6
    top, bot = fake_band(4, w, 0, 1), fake_band(4, w, 1, 0)   # top: rows 0..2 owned + ghost; bot: ghost + rows 1..3
    # a flat piece in `top` whose root is cell (2,1); it touches the seam at columns 1..3 (edge row 2) and
    # continues into the ghost row 3 (same root, the local union-find links them)
    top.root[2, 1:4] = top.root[2, 1].item()
    top.root[3, 1:4] = top.root[2, 1].item()
    # the same physical cells seen from `bot`: ghost row 0 and edge row 1, its own root (1,1)
    bot.root[0, 1:4] = bot.root[1, 1].item()
    bot.root[1, 1:4] = bot.root[1, 1].item()
    bot.rootflag[bot.root[1, 1].item()] = 1                  # only the lower piece has seen an outlet
    assert top.merge_flags(1, bot.flag_payload(0)) is True
    assert top.rootflag[top.root[2, 1].item()] == 1
    assert top.merge_flags(1, bot.flag_payload(0)) is False  # stable now
    assert bot.merge_flags(0, top.flag_payload(1)) is False  # nothing new for the lower band
    # NoData seam cells are ignored
    top2 = fake_band(4, w, 0, 1)
    top2.ft[2, :] = CudaFlatsBand.FT_NODATA
    top2.ft[3, :] = CudaFlatsBand.FT_NODATA
    assert top2.merge_flags(1, torch.ones((2, w), dtype=torch.uint8)) is False


def test_heights_take_the_maximum_across_a_seam():
    w = 5
    top, bot = fake_band(3, w, 0, 1), fake_band(3, w, 1, 0)
    # labels (root + 1): one flat touching the seam in columns 0..2
    top.root[:] = 0
    bot.root[:] = 0
    top.root[1, 0:3] = 8       # label 8 -> height slot 7  (edge row 1)
    top.root[2, 0:3] = 8       # ghost row
    bot.root[0, 0:3] = 3       # label 3 -> height slot 2  (ghost row)
    bot.root[1, 0:3] = 3       # edge row
    top.height[7] = 5
    bot.height[2] = 9
    assert top.merge_heights(1, bot.height_payload(0)) is True and int(top.height[7]) == 9
    assert bot.merge_heights(0, top.height_payload(1)) is False and int(bot.height[2]) == 9
    assert top.merge_heights(1, bot.height_payload(0)) is False
