"""Host-side proof obligation of the density pass (csrc/sph_device.cuh for_each_window_slot): the
x-window a particle scans, bins [bin(x - h'), bin(x + h')] with h' = h (1 + 1e-6) clipped to its three cells (and
widened to its own bin when it is clamped), must contain the x-bin of EVERY neighbour the reference's fp32 predicate
accepts.  The device formula is restated in numpy (fp64, same operation order) and checked against the oracle's
neighbour sets on lattice, evolved and deliberately awkward states; no GPU involved."""
import numpy as np
import pytest

from oracle_binding import Oracle

H = np.float32(0.0457)
XB = 4


def x_bins(x, box_x, rx):
    """(clamped bin, unclamped window lo, unclamped window hi) per particle, as the device computes them."""
    hb = np.float64(np.float32(box_x)) / 2.0
    h_d = np.float64(H)
    h_win = h_d * (1.0 + 1e-6)
    xd = x.astype(np.float64)

    def unclamped(v):
        return np.floor(((v + hb) / h_d) * XB).astype(np.int64)

    own = np.clip(unclamped(xd), 0, rx * XB - 1)
    return own, unclamped(xd - h_win), unclamped(xd + h_win)


def check_state(o, box):
    bx = box if np.isscalar(box) else box[0]
    rx = o.grid_res[0]
    pos = o.pos
    o.update_grid()
    o.update_density_pressure()
    counts, flat = o.neighbours()
    own, lo_u, hi_u = x_bins(pos[:, 0], bx, rx)
    cx = own // XB
    xl, xr = np.maximum(cx - 1, 0), np.minimum(cx + 1, rx - 1)
    lo = np.minimum(np.maximum(lo_u, xl * XB), own)
    hi = np.maximum(np.minimum(hi_u, (xr + 1) * XB - 1), own)
    i_of = np.repeat(np.arange(o.n), counts)
    bj = own[flat]
    bad = (bj < lo[i_of]) | (bj > hi[i_of])
    assert not bad.any(), f"{int(bad.sum())} neighbours outside their particle's x-window"
    # the window is not vacuous: on average it is clearly narrower than the three cells
    return float(np.mean(hi - lo + 1)) / (3 * XB)


@pytest.mark.parametrize("box,steps", [(0.4, 0), (0.4, 30), (0.9, 0), (0.9, 12), ((0.5, 0.3, 0.4), 20)])
def test_window_contains_every_neighbour_dam_break(box, steps):
    o = Oracle(box).setup_scene()
    if steps:
        o.step(steps)
    frac = check_state(o, box)
    assert frac < 0.9


def test_window_contains_every_neighbour_awkward_states():
    rng = np.random.default_rng(0xC0FFEE)
    box = 0.5
    for case in range(6):
        n = 1500
        pos = (rng.random((n, 3)) - 0.5) * box
        if case == 1:    # dense clump
            pos = (rng.random((n, 3)) - 0.5) * 0.05
        elif case == 2:  # particles outside the box on every side (clamped into the edge cells)
            pos *= 1.3
        elif case == 3:  # exactly on cell faces and bin boundaries
            k = rng.integers(-5, 6, size=(n, 3))
            pos = (k * (float(H) / 4)).astype(np.float64)
        elif case == 4:  # pairs at distance ~h along x (knife edge of the predicate and of the window)
            base = (rng.random((n // 2, 3)) - 0.5) * box * 0.8
            eps = rng.integers(-3, 4, size=n // 2) * np.spacing(np.float32(H))
            other = base.copy()
            other[:, 0] += float(H) + eps
            pos = np.concatenate([base, other])
        elif case == 5:  # everything in one y-z column, spread along x
            pos[:, 1:] *= 0.02
        o = Oracle(box)
        o.set_state(pos.astype(np.float32), np.zeros((len(pos), 3), np.float32))
        check_state(o, box)


def test_fine_key_refines_the_reference_cell_id():
    """x_bin // xb must be the reference's clamped cell coordinate bit for bit (q * xb is exact for a power of two),
    so that a reference cell is one contiguous range of the finer sort key (csrc/sph_device.cuh cell_key)."""
    rng = np.random.default_rng(7)
    for box in (0.4, 0.9, 3.62, 4.56):
        rx = int(np.ceil(np.float32(box) / H))
        hb = np.float64(np.float32(box)) / 2.0
        x = ((rng.random(200000) - 0.5) * box * 1.2).astype(np.float32)
        # add the exact cell faces and their fp32 neighbours
        faces = (np.arange(-2, rx + 3) * np.float64(H) - hb).astype(np.float32)
        x = np.concatenate([x, faces, np.nextafter(faces, np.float32(np.inf)), np.nextafter(faces, np.float32(-np.inf))])
        q = (x.astype(np.float64) + hb) / np.float64(H)
        ref_cell = np.clip(np.floor(q).astype(np.int64), 0, rx - 1)          # src/CCPUParticleSimulator.cpp:46-70
        for xb in (1, 2, 4, 8):
            fine = np.clip(np.floor(q * xb).astype(np.int64), 0, rx * xb - 1)
            assert np.array_equal(fine // xb, ref_cell)
