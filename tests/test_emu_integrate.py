"""CPU: the integrate and frame-pack KERNEL SOURCES (csrc/vh_integrate.cu), compiled for the host over the SIMT stand-in in
tests/emu and executed lane by lane, against the oracle — bit for bit.

This is what lets the kernel be checked where no GPU exists: the same expressions, guard bands, lane mapping, block
scheduler, whole-block discard and negative-voxel bookkeeping as on the device; only MUFU.RCP (emulated as the correctly
rounded reciprocal moved by up to one ulp) and the memory system differ. The GPU parity tests (-m gpu) remain the
authority for the compiled SASS.
"""
import numpy as np
import pytest

from emu.binding import EmuMap


def run_case(ob, synth, *, scene_kw, vox, trunc, maxd=3.0, frames=3, color=True, mutate=None, **kernel_kw):
    sc = synth.Scene(**scene_kw)
    o = ob.Oracle(ob.params_for_scene(sc, voxels_per_block=8, vox_size=vox, trunc_margin=trunc, max_depth=maxd))
    m = EmuMap(sc, vox, trunc, maxd, pool_blocks=1 << 13, color=color)
    total_updates = total_culled = 0
    max_weight = 0.0
    run_case.slow_steps = 0
    for i in range(frames):
        d, rgb, c2w = sc.frame(i)
        if mutate is not None:
            d = mutate(i, d)
        o.begin_frame(c2w)
        o.stage_allocate(d)
        keys = o.visible_keys()
        upd = o.stage_integrate(d, rgb if color else None)
        io = m.integrate(keys, d, rgb, c2w, rcp_seed=17 + i, **kernel_kw)
        assert io.engine_error == 0
        assert io.voxel_updates == upd, f"frame {i}: {io.voxel_updates} voxel updates, oracle {upd}"
        assert io.mismatch == 0
        total_updates += upd
        total_culled += io.culled
        run_case.slow_steps += io.slow_steps
        allk = o.all_keys()
        so, wo, co, found = o.get_blocks(keys)
        se, we, ce, slots = m.blocks(keys)
        assert np.array_equal(se.view(np.uint32), so.view(np.uint32)), f"frame {i}: sdf differs"
        assert np.array_equal(we.view(np.uint32), wo.view(np.uint32)), f"frame {i}: weight differs"
        if color:
            assert np.array_equal(ce[..., :3], co), f"frame {i}: colour differs"
            assert not ce[..., 3].any()
        assert np.array_equal(m.neg[slots], (se < 0).sum(1)), f"frame {i}: negative-voxel counters differ"
        assert len(allk) >= len(keys)
        max_weight = max(max_weight, float(we.max(initial=0)))
    assert max_weight >= min(frames, 2), "the frames of this case do not overlap: the running average is not exercised"
    return total_updates, total_culled


SMALL = dict(width=160, height=120, room=(4.0, 3.0, 2.5), n_frames=60, spheres=((2.8, 1.5, 1.0, 0.4),), color=True)


# integrate kernels: 1 = integrate_kernel_direct (per-lane plane loads), 2 = integrate_kernel_staged (planes staged in shared memory by
# bulk copies), 3 = the direct kernel with 64-bit voxel indices; all consume the work list of cull_list_kernel
REVS = [1, 2, 3]


@pytest.mark.parametrize("rev", REVS)
def test_emulated_integrate_matches_oracle(ob, synth, rev):
    upd, culled = run_case(ob, synth, scene_kw=SMALL, vox=0.04, trunc=0.2, variant=rev)
    assert upd > 20000


@pytest.mark.parametrize("rev", REVS)
def test_emulated_integrate_fine_voxels_discards_blocks(ob, synth, rev):
    """small truncation band: most visible blocks lie behind the surface and the whole-block discard must drop them
    without changing a voxel; the same frames with the discard off give the same map."""
    kw = dict(scene_kw=dict(SMALL, holes=0.02), vox=0.02, trunc=0.06, frames=2, variant=rev)
    upd, culled = run_case(ob, synth, **kw)
    assert culled > 0
    upd2, culled2 = run_case(ob, synth, cull=0, **kw)
    assert culled2 == 0 and upd2 == upd


@pytest.mark.parametrize("kernel_kw", [dict(variant=1, two_steps=1), dict(variant=1, ctas=1),
                                       dict(variant=1, exact_color=1), dict(variant=1, verify=1), dict(variant=1, verify=1, exact_color=1),
                                       dict(variant=2, two_steps=1), dict(variant=2, exact_color=1), dict(variant=2, verify=1, two_steps=1), dict(variant=2, ctas=1)])
def test_emulated_integrate_variants(ob, synth, kernel_kw):
    run_case(ob, synth, scene_kw=SMALL, vox=0.05, trunc=0.2, frames=2, **kernel_kw)


@pytest.mark.parametrize("rev", REVS)
def test_emulated_integrate_no_colour_and_negative_coordinates(ob, synth, rev):
    sc = dict(width=160, height=120, room=(4.0, 3.0, 2.5), room_min=(-2.0, -1.5, -1.25), n_frames=60)
    run_case(ob, synth, scene_kw=sc, vox=0.04, trunc=0.2, frames=2, color=False, variant=rev)


@pytest.mark.parametrize("rev", REVS)
def test_emulated_integrate_hostile_depth(ob, synth, rev):
    """NaN, +-inf, negative, denormal and beyond-MaxDepth samples go through the same gates as in the reference"""
    def mutate(i, d):
        d = d.copy()
        rng = np.random.RandomState(5 + i)
        bad = np.array([np.nan, np.inf, -np.inf, -1.0, 1e-40, 50.0, 0.0], np.float32)
        idx = rng.randint(0, d.size, 600)
        d.reshape(-1)[idx] = bad[rng.randint(0, len(bad), 600)]
        return d
    run_case(ob, synth, scene_kw=SMALL, vox=0.05, trunc=0.2, frames=2, mutate=mutate, variant=rev)
    if rev >= 1:      # NaN / inf numerators leave the fast path: the out-of-line IEEE redo of a step is exercised
        assert run_case.slow_steps > 0
