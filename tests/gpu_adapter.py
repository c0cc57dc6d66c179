"""Gives tractor3d_b200.ParticleEmitter (the CUDA backend behind the C ABI) the same small interface as
cpu_harness.CpuEmitter so one scenario script drives the reference, the oracle and the GPU alike."""
import numpy as np

from tractor3d_b200 import abi
from tractor3d_b200.emitter import Context, ParticleEmitter


class GpuLib:
    """Stands where a CpuLib would: `emitter(params)` creates an emitter on the context."""

    kind = "gpu"

    def __init__(self, rng_mode=abi.RNG_STREAM, seed=0, device=0, rot_mode=abi.ROT_FROZEN):
        self.ctx = Context(device)
        # the rate-cap accumulator and the latched rotation live once per process (like the reference's statics): a test
        # that stands for "a fresh process" starts from fresh ones
        self.ctx.reset_statics()
        self.ctx.set_rotation_mode(rot_mode)
        self.rng_mode, self.seed = rng_mode, seed

    def emitter(self, params):
        return GpuAdapter(ParticleEmitter.create(self.ctx, params, None, (self.rng_mode, self.seed)))

    def close(self):
        self.ctx.close()


class GpuAdapter:
    def __init__(self, pe):
        self.pe = pe

    def set_sprite_frame_coords(self, n, w, h):
        self.pe.setSpriteFrameCoords(n, w, h)

    def set_sprite_tex_coords(self, c):
        self.pe.setSpriteTexCoords(c)

    def set_sprite_frame_rects(self, r):
        self.pe.setSpriteFrameRects(r)

    def sprite_tex_coords(self):
        return self.pe.sprite_tex_coords()

    def set_emission_rate(self, r):
        self.pe.setEmissionRate(r)

    def set_params(self, p):
        self.pe.set_params(p)

    def params(self):
        return self.pe.params()

    def set_node_world(self, m):
        self.pe.set_node_world(m)

    def set_camera_world(self, m):
        self.pe.set_camera_world(m)

    def feed_draws(self, d):
        self.pe.feed_draws(d)

    def start(self):
        self.pe.start()

    def stop(self):
        self.pe.stop()

    def is_active(self):
        return self.pe.isActive()

    def is_started(self):
        return self.pe.isStarted()

    def emit_once(self, n):
        self.pe.emitOnce(n)

    def update(self, ms):
        self.pe.update(ms)

    def draw(self):
        return self.pe.draw()

    def count(self):
        return self.pe.getParticlesCount()

    def state(self):
        return self.pe.state()

    def vertices(self):
        return self.pe.vertices()

    def close(self):
        self.pe.release()
