"""Bank-conflict model of the shared-memory layouts of warp_block_diff and the sigma-14 blur H pass
(guetzli-cuda-opencl_b200/csrc/gzb_device_math.cuh, gzb_kernels.cuh). Pure index arithmetic, no GPU: it
restates the address formulas of the kernels and checks the properties the layouts were chosen for
(profiles/r1_ab_shared_conflicts.md has the measured effect). 32 banks of 4 bytes; a 64-bit access is
served per half-warp, a 128-bit access per quarter-warp; lanes reading the same address do not conflict.
"""
K_BD_PLANE, K_BD_SPEC = 72, 56   # doubles: plane stride of the 8x9 planes / of the row spectra


def wavefronts(addr_by_lane, bytes_per_access):
    """addr_by_lane: {lane: address in units of the access size}. Wavefronts the request needs."""
    words = bytes_per_access // 4
    group = {4: 32, 8: 16, 16: 8}[bytes_per_access]
    total = 0
    for g0 in range(0, 32, group):
        per_bank = {}
        for lane, a in addr_by_lane.items():
            if g0 <= lane < g0 + group:
                per_bank.setdefault((a * words) % 32, set()).add(a)
        if per_bank:
            total += max(len(v) for v in per_bank.values())
    return total


def groups_used(addr_by_lane, bytes_per_access):
    group = {4: 32, 8: 16, 16: 8}[bytes_per_access]
    return len({lane // group for lane in addr_by_lane})


def test_row_transform_reads_are_conflict_free():
    # step (3): lane = plane * 8 + row reads pl[72 * plane + 9 * row + k]
    for k in range(8):
        acc = {lane: K_BD_PLANE * (lane >> 3) + 9 * (lane & 7) + k for lane in range(32)}
        assert wavefronts(acc, 8) == 2   # one per half-warp


def column_task(lane, split):
    """(plane, u) of the column transform a lane runs, or None. split = the half-warp mapping."""
    if split:
        hl = lane & 15
        return (2 * (lane >> 4) + (1 if hl >= 5 else 0), hl - 5 if hl >= 5 else hl) if hl < 10 else None
    return (lane // 5, lane % 5) if lane < 20 else None


def test_column_transform_reads_and_power_stores():
    for split, swizzle, want_reads, want_stores in ((False, False, 3, None), (True, True, 2, 2)):
        tasks = {lane: column_task(lane, split) for lane in range(32) if column_task(lane, split)}
        assert sorted(tasks.values()) == [(p, u) for p in range(4) for u in range(5)]   # every task exactly once
        for k in range(8):
            reads = {lane: K_BD_SPEC * p + 9 * u + k for lane, (p, u) in tasks.items()}
            assert wavefronts(reads, 8) == want_reads
        for v in range(8):
            stores = {lane: K_BD_PLANE * p + 8 * u + ((v ^ u) if swizzle else v) for lane, (p, u) in tasks.items()}
            n = wavefronts(stores, 8)
            if want_stores is None:
                assert n >= 8     # the old layout: 20 lanes on two bank pairs
            else:
                assert n == want_stores == groups_used(stores, 8)


def test_swizzled_power_rows_are_read_back_where_they_were_stored():
    # step (5) reads bin i = 8u + v (4 <= i <= 36) at i ^ (i >> 3); step (4) stored it at 8u + (v ^ u)
    for i in range(4, 37):
        u, v = i >> 3, i & 7
        assert i ^ (i >> 3) == 8 * u + (v ^ u)
    # Lanes 0..15 (bins 4..19) still read 16 consecutive doubles. Lanes 16..31 read bins 20..35: row 4's
    # v = 0..3 are swizzled to slots 4..7, the banks of row 2's slots 4..7 -> one extra wavefront per read (four
    # reads per cell against eight stores that lost two to three wavefronts each). A row permutation
    # f = (0, 1, 2, 4, 3) instead of the identity (slot = v ^ f(u)) would keep both conflict-free; not measured yet.
    acc = {lane: (4 + lane) ^ ((4 + lane) >> 3) for lane in range(32)}
    assert wavefronts({l: a for l, a in acc.items() if l < 16}, 8) == 1
    assert wavefronts(acc, 8) == 3
    f = (0, 1, 2, 4, 3)
    alt = {lane: ((4 + lane) & ~7) | (((4 + lane) & 7) ^ f[(4 + lane) >> 3]) for lane in range(32)}
    assert wavefronts(alt, 8) == 2
    tasks = {lane: column_task(lane, True) for lane in range(32) if column_task(lane, True)}
    for v in range(8):
        assert wavefronts({lane: K_BD_PLANE * p + 8 * u + (v ^ f[u]) for lane, (p, u) in tasks.items()}, 8) == 2


def test_sigma14_h_pass_chunk_reads():
    # outputs four pixels apart: scalar reads of tap k put the warp on 8 banks, 16-byte chunks do not
    for off in range(4):
        scalar = {lane: off + 4 * lane + 7 for lane in range(32)}          # any tap index
        assert wavefronts(scalar, 4) == 4
        chunks = {lane: lane + 3 for lane in range(32)}                     # chunk m of lane: (4 * lane) / 4 + m
        assert wavefronts(chunks, 16) == 4                                  # 4 quarter-warps x 1: four taps per lane
    # the last chunk a lane touches stays inside the TMA box: box_w = (span + 6) & ~3, span = 4 * 31 + nt
    for r in (31, 32):
        nt = 2 * r + 1
        box_w = (4 * 31 + nt + 6) & ~3
        for off in range(4):
            nchunks = (off + nt + 3) // 4
            assert 4 * 31 + 4 * nchunks <= box_w
