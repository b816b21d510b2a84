"""Host-side logic: mesh readers/generators and the case-file format (CPU)."""
import numpy as np

from gpupfem2_b200 import casefile
from gpupfem2_b200.mesh import load_dat, poiseuille_field, structured_channel


def test_load_dat_format(tmp_path):
    # reference format (src/mesh_2d.cu:36-96): header, vertices "id x y z", entities "id type ...", 203 = triangle
    p = tmp_path / "m.dat"
    p.write_text("4 4\n1 0.0 0.0 0.0\n2 1.0 0.0 0.0\n3 1.0 1.0 0.0\n4 0.0 1.0 0.0\n"
                 "1 102 1 2 \n2 102 2 3 \n3 203 1 2 3 \n4 203 1 3 4 \n")
    m = load_dat(str(p), scale=2.0)
    assert m.n_nodes == 4 and m.n_cells == 2
    assert np.array_equal(m.cells, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32))  # 1-based -> 0-based
    assert np.array_equal(m.vertices[2], [2.0, 2.0])


def test_structured_channel_numbering():
    for colmajor in (True, False):
        m = structured_channel(5, 3, 10.0, 3.0, colmajor=colmajor)
        assert m.n_nodes == 6 * 4 and m.n_cells == 2 * 5 * 3
        v, c = m.vertices, m.cells.astype(np.int64)
        a, b, d = v[c[:, 0]], v[c[:, 1]], v[c[:, 2]]
        area = 0.5 * ((b[:, 0] - a[:, 0]) * (d[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (d[:, 0] - a[:, 0]))
        assert np.allclose(area, 0.5 * 2.0 * 1.0) and np.all(area > 0)  # counter-clockwise
        assert np.isclose(area.sum(), 30.0)
        if colmajor:  # a strip in x is a contiguous cell range: centroid x is non-decreasing in blocks of 2*ny
            cx = (a[:, 0] + b[:, 0] + d[:, 0]) / 3
            blocks = cx.reshape(5, 6)
            assert np.all(blocks.max(axis=1)[:-1] < blocks.min(axis=1)[1:])


def test_poiseuille_field():
    m = structured_channel(2, 4, 1.0, 2.0)
    fx, fy = poiseuille_field(m, 3.0, 2.0)
    assert np.all(fy == 0)
    assert np.isclose(fx.max(), 3.0) and fx.min() == 0.0


def test_casefile_roundtrip(tmp_path, oracle):
    m = oracle.complete_mesh(structured_channel(3, 2, 1.0, 1.0))
    fx, fy = poiseuille_field(m, 1.0, 1.0)
    p = tmp_path / "c.bin"
    casefile.write_case(str(p), m, fx, fy, 2, 3, 0.01, 5, (1, 5))
    raw = np.fromfile(str(p), dtype=np.uint8)
    hdr = raw[:64].view(np.int64)
    assert hdr[0] == casefile.MAGIC and hdr[1] == m.n_nodes and hdr[2] == m.n_cells and hdr[7] == 2
    expect = 64 + 8 + 16 + m.n_nodes * 16 + m.n_cells * 12 + (m.n_cells + 1) * 4 + m.nbr_indices.size * 4 + m.n_nodes * 16
    assert raw.size == expect


def test_lazy_resort_index_model():
    """Index-level model of the lazy re-sort (gpupfem2_b200/csrc/pfem2_lazy.cuh + advect_lazy in pfem2_api.cu): the records move once per
    step, a permutation src[] stands in for the physical sort.  The model mirrors the kernels' index logic (not their arithmetic) on random
    movement with deletions and re-seeding and checks what the CUDA path relies on: src_new is a bijection onto the live rows, the
    segments are sorted by cell, the dense array plus the appended re-seeds never exceeds n_old + added rows, the index arrays are padded
    to whole tiles with valid rows, the chunked move pass partitions the tiles exactly, and materialise gives the order a physical
    counting sort gives (as a set per cell: the slot order inside a cell is scheduling-dependent on the GPU)."""
    rng = np.random.default_rng(5)
    C, ppc, LOST = 97, 4, 0xFFFFFFFF
    # state: two record buffers (payload = a unique tag per particle + its cell), cur, the permutation, the segment table
    cap = 4096
    tag = [np.zeros(cap, np.int64), np.zeros(cap, np.int64)]
    cell = [np.full(cap, LOST, np.int64), np.full(cap, LOST, np.int64)]
    cur, count, next_tag = 0, C * ppc, C * ppc
    tag[0][:count] = np.arange(count)
    cell[0][:count] = np.repeat(np.arange(C), ppc)
    cell_start = np.arange(C + 1) * ppc
    src, permuted = None, False
    for step in range(12):
        if not permuted:  # k_iota: identity, padded to a whole tile with row 0
            padded = (count + 31) & ~31
            src = np.where(np.arange(padded) < count, np.arange(padded), 0)
        n_old = count
        assert src.shape[0] == (n_old + 31) & ~31 and np.all(src < cap)
        # chunked move pass (k_advect_locate_lazy with chunk_start): whole tiles, exact partition of [0, tiles)
        K = 1 + step % 4
        cb = [C * j // K for j in range(K + 1)]
        covered = np.zeros((n_old + 31) // 32, np.int64)
        out_tag, out_cell = tag[cur ^ 1], cell[cur ^ 1]
        keys = np.full(((n_old + 31) & ~31), LOST, np.int64)
        for j in range(K):
            p_lo = min(cell_start[cb[j]], n_old) & ~31
            n_hi = n_old if cb[j + 1] >= C else (min(cell_start[cb[j + 1]], n_old) & ~31)
            for base in range(p_lo, n_hi, 32):
                covered[base // 32] += 1
                rows = src[base:base + 32]                      # 8 x gather4
                t, c = tag[cur][rows].copy(), cell[cur][rows].copy()
                valid = base + np.arange(32) < n_old
                move = rng.integers(-3, 4, 32)                  # new cell: a neighbour, or lost
                newc = np.where(rng.random(32) < 0.05, LOST, np.clip(c + move, 0, C - 1))
                newc = np.where(valid, newc, LOST)
                out_tag[base:base + 32] = t                     # dense tile store (padding lanes: a stale copy of row src = 0)
                out_cell[base:base + 32] = np.where(valid, newc, c)
                keys[base:base + 32] = newc
        assert np.all(covered == 1), "the chunks do not partition the tiles"
        cur ^= 1
        # k_plan_cells / scan: every cell is topped up to at least ppc particles (the model's stand-in for the sub-cell check)
        live = np.bincount(keys[keys != LOST], minlength=C)
        missing = np.maximum(ppc - live, 0)
        start = np.concatenate([[0], np.cumsum(live + missing)])
        count = int(start[-1])
        # k_rank: slot = cursor[cell]++ in an arbitrary (here: shuffled) order of the records
        src_new = np.full((count + 31) & ~31, -1, np.int64)
        cursor = start[:-1].copy()
        for i in rng.permutation(n_old):
            if keys[i] != LOST:
                src_new[cursor[keys[i]]] = i
                cursor[keys[i]] += 1
        # k_reseed_lazy: appended rows behind the dense array, indices behind the cell's survivors; pad behind the last position
        tail = 0
        for c in rng.permutation(C):
            for k in range(missing[c]):
                d = n_old + tail
                tail += 1
                assert d < cap
                tag[cur][d], cell[cur][d] = next_tag, c
                next_tag += 1
                src_new[start[c] + live[c] + k] = d
        src_new[count:] = 0
        assert tail == missing.sum() and n_old + tail <= cap
        # invariants the projection / the next move pass rely on
        s = src_new[:count]
        assert np.all(s >= 0) and np.unique(s).shape[0] == count, "src_new is not a bijection onto the live rows"
        assert np.all(np.diff(cell[cur][s]) >= 0) and np.all(cell[cur][s] != LOST)
        assert np.array_equal(np.bincount(cell[cur][s], minlength=C), np.diff(start))
        src, cell_start, permuted = src_new, start, True
        if step % 5 == 4:  # k_materialize: out[j] = in[src[j]] -> the physically sorted array
            tag[cur ^ 1][:count] = tag[cur][s]
            cell[cur ^ 1][:count] = cell[cur][s]
            cur ^= 1
            permuted = False
            assert np.all(np.diff(cell[cur][:count]) >= 0)
            assert np.unique(tag[cur][:count]).shape[0] == count


def test_partitioned_mesh_slices_find_the_global_interface_nodes():
    """Mesh partition (multi_gpu.channel_slice_columns / the `cell_base` convention): every rank only holds its own quad
    columns + the halo, in a numbering of its own.  The interface node lists it derives from the slice must be the global
    ones (shifted by the slice's first node) in the same ascending order on both sides of every boundary, and the halo must
    cover band x (substeps + 1) cells."""
    import torch

    from gpupfem2_b200 import multi_gpu
    from gpupfem2_b200.mesh import structured_channel

    nx, ny, world, S = 36, 5, 4, 3
    m = structured_channel(nx, ny, 3.6, 0.5, colmajor=True)
    cells = m.cells.astype(np.int64)
    bounds = multi_gpu.strip_bounds(m.n_cells, world, align=2 * ny)
    band = 2 * ny + 3  # vertex-sharing one-ring of the x-major channel (checked against the device value in the GPU tests)
    assert multi_gpu.halo_cells(band, S) == band * (S + 1)
    glob = {r: multi_gpu.interface_nodes(torch.as_tensor(cells), bounds, r) for r in range(world)}
    for r in range(world):
        c0, c1 = multi_gpu.channel_slice_columns(nx, ny, bounds, r, band, S)
        lo, hi = 2 * ny * c0, 2 * ny * c1
        assert lo <= max(0, bounds[r] - multi_gpu.halo_cells(band, S)) and hi >= min(m.n_cells, bounds[r + 1] + multi_gpu.halo_cells(band, S))
        node_base = (ny + 1) * c0
        local_cells = cells[lo:hi] - node_base
        assert local_cells.min() == 0 and local_cells.max() == (ny + 1) * (c1 - c0 + 1) - 1  # exactly the slice's own nodes
        local_bounds = np.clip(bounds.astype(np.int64) - lo, 0, hi - lo)
        mine = multi_gpu.interface_nodes(torch.as_tensor(local_cells), local_bounds, r)
        assert sorted(mine) == sorted(glob[r])
        for nb, idx in mine.items():
            assert abs(nb - r) == 1
            assert torch.equal(idx + node_base, glob[r][nb])


def test_tile_claim_index_model():
    """Index-level model of how the gathered move pass hands out its tiles (pfem2_move.cuh, k_move_gather with PFEM2_MOVE_GDYN): every
    warp starts with the fixed group `global warp`, keeps two pending positions and one raw claim, and turns to a claimed group when it
    has taken the last tile of the group it is in.  Whatever the order in which the warps advance, every tile of [p_lo, n) must be
    taken exactly once and a claim must be at least two iterations old when its first tile becomes the current one."""
    rng = np.random.default_rng(5)
    for claim in (1, 2, 4, 8):
        for warps, tiles, p_lo in ((8, 0, 0), (8, 5, 64), (8, 8 * claim, 0), (24, 1000, 32), (16, 777, 0), (4, 64 * claim + 3, 96)):
            stride = warps * 32
            n = p_lo + tiles * 32 - (7 if tiles else 0)  # a partial last tile
            last = (claim - 1) << 5
            dyn0 = p_lo + stride * claim
            cursor = [0]

            def atomic_add(v):
                r = cursor[0]
                cursor[0] += v
                return r

            def follow(pos, c):
                return pos + 32 if ((pos - p_lo) & last) != last else dyn0 + c * (claim * 32)

            class Warp:
                def __init__(self, w):
                    self.base = p_lo + ((w * claim) << 5)
                    self.pending = atomic_add(3 if claim == 1 else (2 if claim == 2 else 1))
                    self.claimed_at = -2  # (prologue claims are used at the earliest two iterations later by construction)
                    self.it = 0
                    p1 = self.base + 32 if claim > 1 else dyn0 + self.pending * 32
                    p2 = self.base + 64 if claim > 2 else (dyn0 + self.pending * 64 if claim == 2 else dyn0 + (self.pending + 1) * 32)
                    if claim <= 2:
                        self.pending += 2 if claim == 1 else 1
                    self.slot = [p1, p2]
                    self.done = self.base >= n

                def step(self, taken):
                    taken.append(self.base)
                    nxt, nn = self.slot
                    self.slot = [nn, follow(nn, self.pending)]
                    if ((nn - p_lo) & last) == last:  # the pending claim was consumed: it must not be younger than one iteration
                        assert self.it - self.claimed_at >= 1
                        self.pending = atomic_add(1)
                        self.claimed_at = self.it
                    self.it += 1
                    self.base = nxt
                    self.done = self.base >= n

            ws = [Warp(w) for w in range(warps)]
            taken = []
            while True:
                live = [w for w in ws if not w.done]
                if not live:
                    break
                live[rng.integers(len(live))].step(taken)  # an arbitrary interleaving of the warps
            want = list(range(p_lo, n, 32))
            assert sorted(taken) == want, (claim, warps, tiles)


def test_reseed_lane_per_particle_index_model():
    """Index-level model of reseed_cells_of_warp (pfem2_resort.cuh): the new particles of the 32 cells of a warp, one lane per new
    particle -- owner lane by the kernel's binary search over the inclusive prefix sums, r-th empty sub-cell of the owner's occupancy
    mask.  Every empty sub-cell of every needy cell must be produced exactly once, in ascending sub-cell order per cell, into the row
    d(owner) + r and the permutation slot j(owner) + r."""
    rng = np.random.default_rng(11)

    def nth_set_bit(word, n):  # __fns(word, 0, n + 1)
        for b in range(32):
            if (word >> b) & 1:
                if n == 0:
                    return b
                n -= 1
        raise AssertionError("rank beyond the set bits")

    for ppc in (1, 4, 16, 36, 64):
        full = (1 << ppc) - 1
        for _ in range(40):
            mask = [int(rng.integers(0, 1 << 62)) & full if rng.random() < 0.6 else full for _ in range(32)]
            missing = [ppc - bin(m).count("1") for m in mask]
            incl = np.cumsum(missing)
            total = int(incl[-1])
            d = [1000 * lane for lane in range(32)]  # first row of the lane's block (any disjoint blocks)
            j = [64 * lane for lane in range(32)]
            got = {}
            for k in range(total):
                o = 0
                step = 16
                while step:  # the kernel's search: smallest o with incl[o] > k
                    if incl[o + step - 1] <= k:
                        o += step
                    step >>= 1
                o = min(o, 31)
                r = k - (int(incl[o]) - missing[o])
                assert 0 <= r < missing[o]
                empty = ~mask[o] & full
                lo, hi = empty & 0xffffffff, empty >> 32
                nlo = bin(lo).count("1")
                s = nth_set_bit(lo, r) if r < nlo else 32 + nth_set_bit(hi, r - nlo)
                assert (o, s) not in got
                got[(o, s)] = (d[o] + r, j[o] + r)
            for lane in range(32):
                subs = [s for s in range(ppc) if not (mask[lane] >> s) & 1]
                assert [got[(lane, s)] for s in subs] == [(d[lane] + r, j[lane] + r) for r in range(len(subs))]
            assert len(got) == total
