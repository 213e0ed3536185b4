"""CPU tests of tools/layout_model.py against layout words dumped from the device (tests/golden/layout_*.npz, written by
tools/dev/dump_lpos.py on a B200 through oar_store_layout_lpos for synth.make_config("small") / ("tiny")):
the CPU restatement of build_tiles' position greedy reproduces the device's x positions exactly, and the invariants the
M-step relies on hold on the dumped tiles."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lm():
    spec = importlib.util.spec_from_file_location("layout_model", os.path.join(ROOT, "tools", "layout_model.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["layout_small_24tiles.npz", "layout_tiny_8tiles.npz"])
def test_emulated_greedy_reproduces_the_device_layout(lm, name):
    d = np.load(os.path.join(ROOT, "tests", "golden", name))
    words, trash = d["words"], d["trash"]
    assert words.shape[1] == 1024 and len(trash) == words.shape[0]
    shipped = earlier = 0
    for t in range(words.shape[0]):
        slots, pos, xd = lm.dump_tile(words[t], trash[t])
        assert lm._xd_of(slots) == xd                                   # the trash slots start right behind the items
        real = pos[slots >= 0]
        items, _ = lm.items_of(slots)
        assert sorted(real.tolist()) == sorted(x + o for its in items.values() for x, n in its for o in range(n))   # perfect assignment
        emu = lm.emulate_build_greedy(slots, xd)                        # the rule as shipped: by supply, scarce first
        assert len(emu) == int((slots >= 0).sum())
        assert all(int(pos[s]) == p for s, p in emu.items()), f"tile {t}: the emulation and the device disagree"
        for g in lm.groups():                                           # one trash slot per half-warp store, in a free bank
            tr = {int(pos[i]) for i in g if slots[i] < 0}
            assert len(tr) <= 1
            used = {int(pos[i]) & 15 for i in g if slots[i] >= 0}
            if tr and len(used) < 16 and any(slots[i] >= 0 for i in g):
                assert (next(iter(tr)) & 15) not in used
        shipped += lm.wavefronts_with_trash(emu, xd, "free")
        earlier += lm.wavefronts_with_trash(lm.emulate_build_greedy(slots, xd, supply=False, premark=True), xd, "lane")
    assert 64 * words.shape[0] <= shipped < earlier                     # 64 half-warp stores per tile is the floor
