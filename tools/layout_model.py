#!/usr/bin/env python
"""CPU model of the chunk layout's shared-memory bank behaviour (no GPU needed).

Restates what build_tiles (oarfish_b200/csrc/oar_tiled.cuh) does to a store -- rows sorted by smallest transcript,
tiles of `span` alignments, first-fit packing into 8 warp-chunks of 127 slots, items of 16/8/4 x slots at strides
18/10/6 doubles -- and counts the wavefronts the sweep's shared-memory instructions need per tile:

  scatter  STS.64 of slot k of the 32 lanes of a warp: per half-warp, max number of positions in one 8-byte bank
  gather   LDS.64 of prev[table index]: per half-warp, max number of DISTINCT table entries in one bank

for several ways of handing out x positions / table indices.  Used to choose the heuristics before spending GPU time.
usage: python tools/layout_model.py [n_reads] [n_txps] [avg] [max_tiles]
"""
import sys
import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from oarfish_b200 import synth

SPAN, CHUNK, CAP, NW, IMAX = 984, 128, 127, 8, 16


def tiles_of(store, max_tiles):
    rp = store.row_ptr.astype(np.int64); tx = store.txp_id
    lens = np.diff(rp)
    mn = np.minimum.reduceat(tx, rp[:-1])
    order = np.argsort(mn, kind="stable")
    off = np.concatenate([[0], np.cumsum(lens[order])])
    n_tiles = int((off[-1] + SPAN - 1) // SPAN)
    starts = np.searchsorted(off[:-1], np.arange(n_tiles + 1) * SPAN, side="left")
    pick = np.linspace(0, n_tiles - 1, min(max_tiles, n_tiles)).astype(int)
    for t in pick:
        rows = order[starts[t]:starts[t + 1]]
        yield [np.sort(tx[rp[r]:rp[r + 1]]) for r in rows]


def pack(rows):
    """slot -> transcript (or -1), first fit; short rows wait for a lane-crossing position (as build_tiles)."""
    used = [0] * NW
    slots = -np.ones(NW * CHUNK, dtype=np.int64)
    pend = []

    def fit(n, clean):
        for c in range(NW):
            if used[c] + n <= CAP and (not clean or (used[c] & 3) + n >= 4):
                return c
        return -1

    def place(r, c):
        slots[c * CHUNK + used[c]: c * CHUNK + used[c] + len(r)] = r
        used[c] += len(r)

    for r in rows:
        if len(r) < 4:
            c = fit(len(r), True)
            if c >= 0: place(r, c)
            else: pend.append(r)
            continue
        c = fit(len(r), False)
        if c < 0: continue
        place(r, c)
        k = 0
        while k < len(pend):
            c2 = fit(len(pend[k]), True)
            if c2 >= 0: place(pend.pop(k), c2)
            else: k += 1
    for r in pend:
        c = fit(len(r), True)
        if c < 0: c = fit(len(r), False)
        if c >= 0: place(r, c)
    return slots


def items_of(slots):
    """per transcript: (count, list of (x offset, n valid) items) in the build's class order."""
    t, cnt = np.unique(slots[slots >= 0], return_counts=True)
    agg = [(int(a), int(b)) for a, b in zip(t, cnt) if b >= 4]
    n0 = sum(c // IMAX + (1 if c % IMAX > IMAX // 2 else 0) for _, c in agg)
    n1 = sum(1 for _, c in agg if IMAX // 4 < c % IMAX <= IMAX // 2)
    i0 = i1 = i2 = 0
    out = {}
    for tr, c in agg:
        q, rem = divmod(c, IMAX)
        its = [((i0 + v) * 18, IMAX) for v in range(q)]
        i0 += q
        if rem > IMAX // 2: its.append((i0 * 18, rem)); i0 += 1
        elif rem > IMAX // 4: its.append((n0 * 18 + i1 * 10, rem)); i1 += 1
        elif rem >= 1: its.append((n0 * 18 + n1 * 10 + i2 * 6, rem)); i2 += 1
        out[tr] = its
    return out, list(t)


def groups():
    for c in range(NW):
        for k in range(4):
            for h in range(2):
                yield [c * CHUNK + 4 * (16 * h + l) + k for l in range(16)]


def scatter_sorted(slots, items):
    """positions in (transcript; chunk, k, lane) sort order -- the layout before the bank-aware assignment."""
    pos = {}
    nxt = {tr: 0 for tr in items}
    flat = {tr: [x + o for x, n in its for o in range(n)] for tr, its in items.items()}
    for g in groups():
        pass
    order = sorted((s for s in range(NW * CHUNK) if slots[s] in items),
                   key=lambda s: (slots[s], s // CHUNK, s % 4, (s % CHUNK) // 4))
    for s in order:
        tr = slots[s]; pos[s] = flat[tr][nxt[tr]]; nxt[tr] += 1
    return pos


def scatter_greedy(slots, items, mode="lane"):
    """the greedy of build_tiles; mode: 'lane' start the residue search at the lane number (current),
    'scarce' = lanes whose transcript offers the fewest residues choose first, then 'lane'."""
    free = {tr: {} for tr in items}          # residue -> list of positions still free
    for tr, its in items.items():
        for x, n in its:
            for o in range(n):
                free[tr].setdefault((x + o) & 15, []).append(x + o)
    pos = {}
    for g in groups():
        G = set()
        lanes = [(l, s) for l, s in enumerate(g) if slots[s] in items]
        if mode == "scarce":
            lanes.sort(key=lambda ls: (len(free[slots[ls[1]]]), ls[0]))
        for l, s in lanes:
            av = free[slots[s]]
            cand = [r for r in av if r not in G] or list(av)
            rho = min(cand, key=lambda r: (r - l) & 15)
            G.add(rho)
            pos[s] = av[rho].pop(0)
            if not av[rho]: del av[rho]
    return pos


def scatter_wavefronts(pos):
    w = 0
    for g in groups():
        banks = {}
        for s in g:
            if s in pos: banks[pos[s] & 15] = banks.get(pos[s] & 15, 0) + 1
        w += max(banks.values()) if banks else 0
    return w


def gather_wavefronts(slots, index):
    w = 0
    for g in groups():
        banks = {}
        for s in g:
            d = index.get(slots[s], 0)
            banks.setdefault(d & 15, set()).add(d)
        w += max(len(v) for v in banks.values())
    return w


def table_coloured(slots, table):
    """table order chosen so that transcripts read by the same half-warp instruction sit in different banks:
    greedy over transcripts by decreasing count, each takes the residue with the least co-occurrence weight."""
    co = {}
    for g in groups():
        ts = {slots[s] for s in g if slots[s] >= 0}
        for a in ts:
            for b in ts:
                if a != b: co.setdefault(a, {}); co[a][b] = co[a].get(b, 0) + 1
    cnt = {t: int((slots == t).sum()) for t in table}
    res_of, load = {}, [0] * 16
    for t in sorted(table, key=lambda t: -cnt[t]):
        cost = [sum(wt for b, wt in co.get(t, {}).items() if res_of.get(b) == r) for r in range(16)]
        r = min(range(16), key=lambda r: (cost[r], load[r]))
        res_of[t] = r; load[r] += 1
    # residue r, j-th transcript with that residue -> index r + 16 j (the table gets holes up to 16 * max load)
    seen = [0] * 16; index = {}
    for t in sorted(table, key=lambda t: -cnt[t]):
        r = res_of[t]; index[t] = r + 16 * seen[r]; seen[r] += 1
    return index, 16 * max(load)


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    n_txps = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
    avg = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
    max_tiles = int(sys.argv[4]) if len(sys.argv) > 4 else 150
    st = synth.make_store(n_reads, n_txps, avg, 3)
    acc = {}
    nt = 0
    for rows in tiles_of(st, max_tiles):
        slots = pack(rows)
        items, table = items_of(slots)
        nt += 1
        for name, pos in (("scatter sorted", scatter_sorted(slots, items)), ("scatter greedy/lane", scatter_greedy(slots, items, "lane")),
                          ("scatter greedy/scarce", scatter_greedy(slots, items, "scarce"))):
            acc[name] = acc.get(name, 0) + scatter_wavefronts(pos)
        acc["gather id order"] = acc.get("gather id order", 0) + gather_wavefronts(slots, {t: i for i, t in enumerate(table)})
        idx, size = table_coloured(slots, table)
        acc["gather coloured"] = acc.get("gather coloured", 0) + gather_wavefronts(slots, idx)
        acc["table size id"] = acc.get("table size id", 0) + len(table)
        acc["table size coloured"] = acc.get("table size coloured", 0) + size
    print(f"{nt} tiles; per tile (64 half-warp instructions each, ideal 64):")
    for k, v in acc.items():
        print(f"  {k:24s} {v / nt:8.1f}")


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("matching", "heuristics", "table", "dump")):
    main()


def scatter_matching(slots, items):
    """per half-warp group a maximum bipartite matching lanes <-> bank residues (augmenting paths) on what the
    transcripts still offer: the best a group-by-group assignment can do (bound for the greedy)."""
    free = {tr: {} for tr in items}
    for tr, its in items.items():
        for x, n in its:
            for o in range(n):
                free[tr].setdefault((x + o) & 15, []).append(x + o)
    pos = {}
    for g in groups():
        lanes = [s for s in g if slots[s] in items]
        # capacities: a transcript offers residue r len(free[tr][r]) times; lanes of one transcript share them
        match_r = {}          # residue -> lane slot
        def try_lane(s, seen):
            tr = slots[s]
            for r in sorted(free[tr], key=lambda r: -len(free[tr][r])):
                if r in seen: continue
                # supply check: lanes of the same transcript already holding r
                held = sum(1 for rr, ss in match_r.items() if rr == r and slots[ss] == tr)
                if held >= len(free[tr][r]): continue
                seen.add(r)
                if r not in match_r or try_lane(match_r[r], seen):
                    match_r[r] = s
                    return True
            return False
        unmatched = []
        for s in sorted(lanes, key=lambda s: len(free[slots[s]])):
            if not try_lane(s, set()): unmatched.append(s)
        for r, s in match_r.items():
            pos[s] = free[slots[s]][r].pop(0)
            if not free[slots[s]][r]: del free[slots[s]][r]
        for s in unmatched:   # conflict: take the residue the transcript has most of
            av = free[slots[s]]
            r = max(av, key=lambda r: len(av[r]))
            pos[s] = av[r].pop(0)
            if not av[r]: del av[r]
    return pos


def main2():
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
    st = synth.make_store(n_reads, n_reads // 50, 8.0, 3)
    acc = {}; nt = 0
    for rows in tiles_of(st, int(sys.argv[3]) if len(sys.argv) > 3 else 60):
        slots = pack(rows); items, table = items_of(slots); nt += 1
        for name, pos in (("scatter greedy/lane", scatter_greedy(slots, items, "lane")), ("scatter greedy/scarce", scatter_greedy(slots, items, "scarce")),
                          ("scatter matching", scatter_matching(slots, items))):
            acc[name] = acc.get(name, 0) + scatter_wavefronts(pos)
    for k, v in acc.items():
        print(f"  {k:24s} {v / nt:8.1f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "matching":
    main2()


def scatter_greedy2(slots, items, augment=True, by_supply=True):
    """scarce-first greedy; constrained lanes take the free residue their transcript has most of; a lane left without a
    free residue may displace one already-placed lane of the group whose transcript can move elsewhere (one-step
    augmentation)."""
    free = {tr: {} for tr in items}
    for tr, its in items.items():
        for x, n in its:
            for o in range(n):
                free[tr].setdefault((x + o) & 15, []).append(x + o)
    pos = {}
    for g in groups():
        lanes = [(l, s) for l, s in enumerate(g) if slots[s] in items]
        lanes.sort(key=lambda ls: (len(free[slots[ls[1]]]), ls[0]))
        taken = {}     # residue -> slot
        held = {}      # (tr, residue) -> count reserved in this group
        def offers(tr, r):
            return r in free[tr] and len(free[tr][r]) > held.get((tr, r), 0)
        todo_conf = []
        for l, s in lanes:
            tr = slots[s]
            cand = [r for r in free[tr] if r not in taken and offers(tr, r)]
            if cand:
                if by_supply and len(free[tr]) < 16:
                    r = max(cand, key=lambda r: (len(free[tr][r]), -((r - l) & 15)))
                else:
                    r = min(cand, key=lambda r: (r - l) & 15)
                taken[r] = s; held[(tr, r)] = held.get((tr, r), 0) + 1
                continue
            moved = False
            if augment:
                for r in list(free[tr]):
                    if not offers(tr, r) or r not in taken: continue
                    s2 = taken[r]; tr2 = slots[s2]
                    alt = [r2 for r2 in free[tr2] if r2 not in taken and offers(tr2, r2)]
                    if alt:
                        r2 = alt[0]
                        held[(tr2, r)] -= 1; taken[r2] = s2; held[(tr2, r2)] = held.get((tr2, r2), 0) + 1
                        taken[r] = s; held[(tr, r)] = held.get((tr, r), 0) + 1
                        moved = True
                        break
            if not moved: todo_conf.append(s)
        for r, s in taken.items():
            tr = slots[s]; pos[s] = free[tr][r].pop(0)
            if not free[tr][r]: del free[tr][r]
        for s in todo_conf:
            av = free[slots[s]]
            r = max(av, key=lambda r: len(av[r]))
            pos[s] = av[r].pop(0)
            if not av[r]: del av[r]
    return pos


def main3():
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
    st = synth.make_store(n_reads, n_reads // 50, 8.0, 3)
    acc = {}; nt = 0
    for rows in tiles_of(st, int(sys.argv[3]) if len(sys.argv) > 3 else 60):
        slots = pack(rows); items, table = items_of(slots); nt += 1
        for name, pos in (("greedy/lane (shipped)", scatter_greedy(slots, items, "lane")),
                          ("scarce first", scatter_greedy(slots, items, "scarce")),
                          ("scarce + by supply", scatter_greedy2(slots, items, augment=False)),
                          ("scarce + by supply + 1-step augment", scatter_greedy2(slots, items, augment=True)),
                          ("scarce + 1-step augment", scatter_greedy2(slots, items, augment=True, by_supply=False)),
                          ("matching", scatter_matching(slots, items))):
            acc[name] = acc.get(name, 0) + scatter_wavefronts(pos)
    for k, v in acc.items():
        print(f"  {k:40s} {v / nt:8.1f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "heuristics":
    main3()


def table_online(slots, table):
    """online colouring: walk the half-warp groups in order; a transcript seen for the first time takes the residue
    that no other transcript of that group holds (ties: the least loaded residue)."""
    res_of, load = {}, [0] * 16
    for g in groups():
        ts = []
        for s in g:
            t = slots[s]
            if t >= 0 and t not in ts: ts.append(t)
        held = {res_of[t] for t in ts if t in res_of}
        for t in ts:
            if t in res_of: continue
            r = min(range(16), key=lambda r: (r in held, load[r]))
            res_of[t] = r; load[r] += 1; held.add(r)
    seen = [0] * 16; index = {}
    for t in table:
        r = res_of.get(t, 0); index[t] = r + 16 * seen[r]; seen[r] += 1
    return index, 16 * max(max(load), 1)


def main4():
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
    st = synth.make_store(n_reads, n_reads // 50, 8.0, 3)
    acc = {}; nt = 0
    for rows in tiles_of(st, int(sys.argv[3]) if len(sys.argv) > 3 else 60):
        slots = pack(rows); items, table = items_of(slots); nt += 1
        acc["gather id order"] = acc.get("gather id order", 0) + gather_wavefronts(slots, {t: i for i, t in enumerate(table)})
        for name, fn in (("gather coloured (co-occurrence)", table_coloured), ("gather coloured (online)", table_online)):
            idx, size = fn(slots, table)
            acc[name] = acc.get(name, 0) + gather_wavefronts(slots, idx)
            acc[name + " table size"] = acc.get(name + " table size", 0) + size
        acc["table size id"] = acc.get("table size id", 0) + len(table)
    for k, v in acc.items():
        print(f"  {k:48s} {v / nt:8.1f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "table":
    main4()


# ---- real tiles: layout words dumped from the device (oar_store_layout_lpos, tools/dev/dump_lpos.py) -----------------
# usage: python tools/layout_model.py dump gpurun_out/lpos_C3.npy [every]
# Recovers each tile's slot -> transcript map and its trash offset XD from the words, counts the scatter wavefronts of
# the layout as built (the half-warp rule was checked against ncu with tools/dev/sts_bank_probe.cu: exact), and
# re-runs the device's round-based greedy (build_tiles) in emulation with its switches, so that a change to the rule
# can be judged on real tiles before any GPU time is spent.

def _xd_of(slots):
    items, _ = items_of(slots)
    n32 = n16 = n8 = 0
    for its in items.values():
        c = sum(n for _, n in its); q, rem = divmod(c, IMAX)
        n32 += q + (1 if rem > IMAX // 2 else 0); n16 += 1 if IMAX // 4 < rem <= IMAX // 2 else 0; n8 += 1 if 1 <= rem <= IMAX // 4 else 0
    return 18 * n32 + 10 * n16 + 6 * n8


def dump_tile(words, xd=None):
    """(slot -> transcript index or -1, slot -> x position, XD) of one tile's 1024 layout words.  XD (the trash offset, in
    doubles) comes with the dump when oar_store_layout_lpos was asked for it; otherwise it is inferred, which can be
    ambiguous for a tile whose trash slots happen to be used once each."""
    pos = (words >> 16).astype(np.int64) // 8
    off = (words & 0xFFFF).astype(np.int64)
    tr = ((off >> 8) << 5) | ((off & 127) >> 2)
    if xd is not None:
        return np.where(pos >= int(xd), -1, tr), pos, int(xd)
    mx = int(pos.max())
    # padding and stray alignments sit in the 16 trash slots behind the items, so XD is one of mx-15 .. mx+1.  Largest first: a
    # smaller value can be consistent too (dropping the whole last item of a transcript that has nothing else gives a shorter,
    # valid-looking tile), a larger one would make shared trash slots look like x slots handed out twice
    for xd in range(mx + 1, max(0, mx - 15) - 1, -1):
        slots = np.where(pos >= xd, -1, tr)
        real = pos[slots >= 0]
        if _xd_of(slots) == xd and len(np.unique(real)) == len(real):
            return slots, pos, xd
    raise ValueError("tile does not parse")


def emulate_build_greedy(slots, xd, supply=True, premark=False, kscarce=8):
    """build_tiles' position greedy: per scatter instruction (chunk, k) rounds of proposals over the 32 lanes, two phases
    (lanes whose transcript offers <= kscarce residues first), one winner per (transcript, residue) and per (half-warp,
    residue): the lowest lane.  supply: prefer the residues the transcript has most slots of; premark: lanes without a
    position occupy bank (xd + lane) & 15 from the start (the per-lane trash slots of an earlier revision)."""
    items, _ = items_of(slots)
    posl = {}
    for t, its in items.items():
        d = {}
        for x, n in its:
            for o in range(n): d.setdefault((x + o) & 15, []).append(x + o)
        posl[t] = d
    out = {}
    for c in range(NW):
        for k in range(4):
            lanes = [c * CHUNK + 4 * l + k for l in range(32)]
            todo = [slots[s] >= 0 and slots[s] in posl for s in lanes]
            G = [0, 0]
            if premark:
                for l in range(32):
                    if not todo[l]: G[l >> 4] |= 1 << ((xd + (l & 15)) & 15)
            for phase in (0, 1):
                while True:
                    props = {}
                    for l in range(32):
                        if not todo[l]: continue
                        t = slots[lanes[l]]
                        av = sum(1 << r for r in posl[t])
                        if phase == 0 and bin(av).count("1") > kscarce: continue
                        cand = av & ~G[l >> 4] or av
                        if supply:
                            top = max(len(v) for v in posl[t].values())
                            best = sum(1 << r for r, v in posl[t].items() if len(v) == top)
                            if cand & best: cand &= best
                        props[l] = (t, min((r for r in range(16) if cand >> r & 1), key=lambda r: (r - (l & 15)) & 15))
                    if not props: break
                    took = [0, 0]
                    for l, (t, rho) in sorted(props.items()):
                        if min(l2 for l2, p in props.items() if p == (t, rho)) != l: continue
                        if min(l2 for l2, p in props.items() if (l2 >> 4) == (l >> 4) and p[1] == rho) != l: continue
                        out[lanes[l]] = posl[t][rho].pop(0)
                        if not posl[t][rho]: del posl[t][rho]
                        todo[l] = False; took[l >> 4] |= 1 << rho
                    G[0] |= took[0]; G[1] |= took[1]
    return out


def wavefronts_with_trash(posd, xd, trash):
    """scatter wavefronts of a tile; trash = 'lane' (slot xd + lane & 15 per lane), 'free' (one slot per half-warp in a bank
    the others leave free) or 'none' (stores of lanes without a position predicated off)"""
    w = 0
    for g in groups():
        banks = {}
        for s in g:
            if s in posd: banks.setdefault(posd[s] & 15, set()).add(posd[s])
        missing = [li for li, s in enumerate(g) if s not in posd]
        if trash == "lane":
            for li in missing: banks.setdefault((xd + li) & 15, set()).add(xd + li)
        elif trash == "free" and missing:
            free = [r for r in range(16) if r not in banks]
            banks.setdefault(free[0] if free else 0, set()).add(-1)
        w += max(len(v) for v in banks.values()) if banks else 0
    return w


def main5():
    words = np.load(sys.argv[2]); every = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    acc = {}; nt = 0
    for t in range(0, words.shape[0], every):
        slots, pos, xd = dump_tile(words[t]); nt += 1
        def add(k, v): acc[k] = acc.get(k, 0) + v
        add("layout as dumped (every store, distinct words per bank)",
            sum(max(len({int(pos[s]) for s in g if int(pos[s]) & 15 == r}) for r in range(16)) for g in groups()))
        for name, kw, tm in (("greedy of round 1/2a: nearest residue, per-lane trash", dict(supply=False, premark=True), "lane"),
                             ("nearest residue, free-bank trash", dict(supply=False), "free"),
                             ("by supply, per-lane trash", dict(premark=True), "lane"),
                             ("by supply, free-bank trash (shipped)", dict(), "free"),
                             ("by supply, trash stores predicated off", dict(), "none")):
            add(name, wavefronts_with_trash(emulate_build_greedy(slots, xd, **kw), xd, tm))
    print(f"{nt} tiles, scatter wavefronts per tile (64 half-warp stores, floor 64):")
    for k, v in acc.items(): print(f"  {k:58s} {v / nt:7.1f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dump":
    main5()
