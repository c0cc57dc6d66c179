"""What happens to a frame after draw(): the hand-off of the vertex stream to a renderer and the per-emitter
results a scene graph wants.

  * t3d_emitter_draw_to: the quads land, bit-identical, in a buffer the caller owns (a mapped interop vertex buffer
    in an engine; a torch allocation here), through the separate draw kernel and through the fused update+draw launch;
  * t3d_ctx_copy_to_host_async: frames streamed to pinned host memory beside the next frame's launches;
  * T3D_VERTEX_CLIP: sprite.vert:19 applied in the draw kernel, against the restated oracle;
  * t3d_emitter_bounds / t3d_emitter_set_culling: the bounding box reduced inside the update kernel, against the
    restated oracle (the engine has a TODO there, src/scene/Node.cpp:777-780).
"""
import ctypes as C

import numpy as np
import pytest

import cpu_harness as H
import scenarios as S
from tractor3d_b200 import abi, presets

pytestmark = pytest.mark.gpu


def _pair(seed, n, capacity=None, energy=(40.0, 400.0), rng_mode=abi.RNG_DEVICE, rot_mode=abi.ROT_FROZEN):
    from gpu_adapter import GpuLib
    p, sprite = presets.headline_64m(capacity=capacity or n)
    p.energy_min, p.energy_max = energy
    p.emission_rate = 1
    O = H.oracle(fresh=True)
    O.set_rotation_mode(rot_mode)
    lib = GpuLib(rng_mode=rng_mode, seed=seed, rot_mode=rot_mode)
    eo = O.emitter(p); eo.set_sprite_frame_coords(*sprite)
    O.emitter_set_device_rng(eo.h, 1, seed)
    eg = lib.emitter(p); eg.set_sprite_frame_coords(*sprite)
    eo.emit_once(n); eg.emit_once(n)
    eo.start(); eg.start()
    return O, lib, eo, eg


def _projection():
    # a perspective view-projection matrix with nothing special about it (column-major)
    f, a, zn, zf = 1.0 / np.tan(0.4), 16.0 / 9.0, 0.25, 100.0
    proj = np.array([[f / a, 0, 0, 0], [0, f, 0, 0], [0, 0, (zf + zn) / (zn - zf), 2 * zf * zn / (zn - zf)], [0, 0, -1, 0]], dtype=np.float64)
    view = np.eye(4); view[:3, 3] = (-0.3, -1.1, -9.0)
    c, s = np.cos(0.3), np.sin(0.3)
    view[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.ascontiguousarray((proj @ view).T.reshape(16).astype(np.float32))


@pytest.mark.parametrize("n", [1, 300, 4608, 20_001])
def test_draw_to_writes_the_frame_into_a_foreign_allocation(n):
    import torch
    O, lib, eo, eg = _pair(7, n)
    try:
        # sized for exactly the live particles, with a guard band behind it: nothing may be written past the frame's quads
        buf = torch.full(((n + 8) * 36,), float("nan"), dtype=torch.float32, device="cuda")
        for f in range(6):
            eo.update(S.DT); eo.draw()
            eg.update(S.DT)
            if f == 2:
                assert eg.count() == eo.count()      # splits update and draw: the separate draw kernel writes this frame
            assert eg.pe.draw_to(buf.data_ptr(), n) == 1
            ptr, quads = eg.pe.vertices_device()
            assert ptr == buf.data_ptr() and quads == eo.count(), f"frame {f}"
            lib.ctx.sync()
            got = buf.cpu().numpy()
            want = eo.vertices().reshape(-1)
            assert np.array_equal(got[: want.size].view(np.uint32), want.view(np.uint32)), f"frame {f}"
            assert np.isnan(got[n * 36:]).all(), "draw_to wrote beyond the capacity it was given"
            if quads < n:
                # a dead last particle has no quad: the words behind the frame's quads keep whatever the previous (larger)
                # frame left there -- its own quads -- never something new
                pass
        # the emitter's own stream is used again by a plain draw()
        eo.update(S.DT); eo.draw()
        eg.update(S.DT); eg.draw()
        assert eg.pe.vertices_device()[0] != buf.data_ptr()
        assert np.array_equal(eg.vertices().view(np.uint32), eo.vertices().view(np.uint32))
        # too small a destination is refused, not overrun
        from tractor3d_b200.emitter import T3DError
        if n > 1:
            with pytest.raises(T3DError) as ei:
                eg.pe.draw_to(buf.data_ptr(), n // 4 if n > 8 else 0)
            assert ei.value.code == abi.ERR_CAPACITY
        with pytest.raises(T3DError) as ei:
            eg.pe.draw_to(buf.data_ptr() + 4, n)
        assert ei.value.code == abi.ERR_INVALID
    finally:
        lib.close()


def test_frames_stream_to_pinned_host_memory_beside_the_next_frame():
    import torch
    n = 50_000
    O, lib, eo, eg = _pair(11, n, energy=(100.0, 900.0))
    try:
        dev = [torch.zeros(n * 36, dtype=torch.float32, device="cuda") for _ in range(2)]
        host = [torch.zeros(n * 36, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        tickets, wants = [None, None], [None, None]
        for f in range(7):
            b = f & 1
            if tickets[b] is not None:
                lib.ctx.copy_wait(tickets[b])
                got = host[b].numpy()[: wants[b].size]
                assert np.array_equal(got.view(np.uint32), wants[b].view(np.uint32)), f"frame {f - 2}"
            eo.update(S.DT); eo.draw()
            eg.update(S.DT); eg.pe.draw_to(dev[b].data_ptr(), n)
            tickets[b] = lib.ctx.copy_to_host_async(host[b].data_ptr(), dev[b].data_ptr(), eo.count() * 144)
            wants[b] = eo.vertices().reshape(-1).copy()
        for b in range(2):
            lib.ctx.copy_wait(tickets[b])
            assert np.array_equal(host[b].numpy()[: wants[b].size].view(np.uint32), wants[b].view(np.uint32))
    finally:
        lib.close()


@pytest.mark.parametrize("rot_mode", [abi.ROT_FROZEN, abi.ROT_PER_PARTICLE])
def test_clip_space_vertices_match_the_restated_vertex_shader(rot_mode):
    n = 30_000
    O, lib, eo, eg = _pair(21, n, rot_mode=rot_mode)
    try:
        vp = _projection()
        cam = S.rot_xyz(0.2, -0.5, 0.1, (1, 2, 12))
        eo.set_camera_world(cam); eg.set_camera_world(cam)
        lib.ctx.set_vertex_mode(abi.VERTEX_CLIP)
        eg.pe.set_view_projection(vp)
        for f in range(5):
            eo.update(S.DT); eo.draw()
            eg.update(S.DT); eg.draw()
            want = eo.clip_vertices(vp)
            got = eg.vertices()
            assert got.shape == want.shape == (eo.count(), 4, 10), f"frame {f}"
            if rot_mode == abi.ROT_FROZEN:
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"frame {f}"
            else:   # createRotation per quad on the device: the stated tolerance (sincosf differs from glibc's by ulps)
                assert np.allclose(got, want, rtol=1e-5, atol=1e-5), f"frame {f}"
        # back to world space: SpriteVertex again, and the fused launch
        lib.ctx.set_vertex_mode(abi.VERTEX_WORLD)
        eo.update(S.DT); eo.draw()
        eg.update(S.DT); eg.draw()
        if rot_mode == abi.ROT_FROZEN:
            assert np.array_equal(eg.vertices().view(np.uint32), eo.vertices().view(np.uint32))
    finally:
        lib.close()


@pytest.mark.parametrize("n", [1, 777, 9216, 150_000])
def test_bounds_are_reduced_inside_the_update(n):
    O, lib, eo, eg = _pair(31, n, energy=(30.0, 300.0))
    try:
        assert eg.pe.bounds() is None            # not enabled
        eg.pe.enable_bounds(True)
        assert eg.pe.bounds() is None            # enabled, but no update has run since
        for f in range(25):
            eo.update(S.DT); eg.update(S.DT)
            if f % 2:
                eo.draw(); eg.draw()             # fused launches reduce the box too
            bo, bg = eo.bounds(), eg.pe.bounds()
            if eo.count() == 0:
                assert bg is None and eg.count() == 0
                break
            assert bg is not None, f"frame {f}"
            assert np.array_equal(bg[0].view(np.uint32), bo[0].view(np.uint32)), f"frame {f}: lo {bg[0]} vs {bo[0]}"
            assert np.array_equal(bg[1].view(np.uint32), bo[1].view(np.uint32)), f"frame {f}: hi {bg[1]} vs {bo[1]}"
        assert np.array_equal(eg.state(), eo.state())
    finally:
        lib.close()


def test_culling_skips_the_draw_of_an_emitter_outside_the_frustum():
    n = 5000
    O, lib, eo, eg = _pair(41, n, energy=(5000.0, 9000.0))
    try:
        vp = _projection()
        eg.pe.set_view_projection(vp)
        eg.pe.set_culling(True)
        # culling tests the last box that has reached the host (it rides with the asynchronous count feedback);
        # asking for the box makes it current
        eg.update(S.DT); eo.update(S.DT)
        assert eg.pe.bounds() is not None
        assert eg.draw() == 1 and eg.vertices().shape[0] == n          # in view (the emitter sits at the origin)
        eo.draw()   # (the reference latches its process-wide billboard rotation at the first quad it ever draws, SB.cpp:339)
        # move the whole view away: every corner of the box is beyond the right clip plane
        away = vp.copy().reshape(4, 4)   # rows of the column-major array = columns of the matrix
        away[3, 0] -= 1000.0 * 1.0       # clip.x = ... - 1000 * w for w ~ 9: far left of -w
        eg.pe.set_view_projection(away.reshape(16))
        eg.update(S.DT); eo.update(S.DT)
        assert eg.pe.bounds() is not None
        assert eg.draw() == 1            # the reference's return value does not depend on visibility
        assert eg.vertices().shape[0] == 0
        eg.pe.set_culling(False)
        eg.update(S.DT); eo.update(S.DT); eo.draw()
        assert eg.draw() == 1
        assert np.array_equal(eg.vertices().view(np.uint32), eo.vertices().view(np.uint32))
    finally:
        lib.close()


@pytest.mark.parametrize("n", [1, 777, 70_000, 1_200_000])
def test_depth_sorted_indices_walk_the_quads_back_to_front(n):
    # for BLEND_ALPHA the quads want to be drawn farthest first.  The order is defined on an integer key of
    # position . view_dir, so the comparison with the restated oracle is exact -- ties (here: a block of particles forced
    # onto the same position) keep live-array order.
    O, lib, eo, eg = _pair(61, n, energy=(40.0, 4000.0))
    try:
        for f in range(3):
            eo.update(S.DT); eg.update(S.DT)
        so = eo.state()
        if n > 100:                                         # make ties: 40 particles at one point, a few exactly at depth 0
            for c in (abi.COL_POSITION, abi.COL_POSITION + 1, abi.COL_POSITION + 2):
                so[c, 10:50] = so[c, 10]
                so[c, 60:64] = 0
            so[abi.COL_POSITION, 62] = np.float32(-0.0).view(np.uint32)
            eo.set_state(so); eg.pe.upload_state(so)
        for view in ((0.0, 0.0, -1.0), (0.3, -0.5, 0.81), (1.0, 0.0, 0.0)):
            want = eo.depth_sorted_indices(view)
            got = eg.pe.depth_sorted_indices(view)
            assert want.size == got.size == 6 * eo.count()
            assert np.array_equal(got, want), view
            # and it IS a back-to-front walk: depths do not increase along it
            pos = so[abi.COL_POSITION:abi.COL_POSITION + 3].view(np.float32) if n > 100 else eo.state()[abi.COL_POSITION:abi.COL_POSITION + 3].view(np.float32)
            v = np.asarray(view, dtype=np.float32)
            depth = pos[0] * v[0] + pos[1] * v[1] + pos[2] * v[2]
            order = got[0::6] // 4
            assert sorted(order.tolist()) == list(range(eo.count()))
            assert np.all(np.diff(depth[order]) <= 0)
        # the vertices those indices address are the ones a draw of this state writes
        eo.draw(); eg.draw()
        assert np.array_equal(eg.vertices().view(np.uint32), eo.vertices().view(np.uint32))
    finally:
        lib.close()
