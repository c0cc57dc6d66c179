"""Emitter parameter sets used by the workloads of BASELINE.json.

`fire`, `smoke` and `explosion` carry the values of the three stock particle descriptions of the
reference's browser sample (samples/browser/res/common/particles/*.particle) and the sizes of their
textures (fire.png 1024x256, smoke.png 64x64, explosion.png 256x256).  Each preset is
(params, sprite) with sprite = (frame_count, frame_width, frame_height) as handed to
setSpriteFrameCoords (ParticleEmitter.cpp:207).
"""
from . import abi


def fire(**kw):
    p = abi.default_params(
        particle_count_max=5000, emission_rate=300, ellipsoid=1,
        orbit_position=1, orbit_velocity=1, orbit_acceleration=0,
        size_start_min=1.5, size_start_max=2.0, size_end_min=0.5, size_end_max=1.0,
        energy_min=750.0, energy_max=1000.0,
        color_start=[1, 0.2, 0, 1], color_start_var=[0.25, 0, 0, 0], color_end=[0, 0, 0, 0], color_end_var=[0, 0, 0, 0],
        position=[0, 0, 0], position_var=[1, 1, 1], velocity=[0, 10, 0], velocity_var=[4, 6, 4],
        acceleration=[0, 2, 0], acceleration_var=[3, 0.5, 3],
        rotation_per_particle_speed_min=-1.5, rotation_per_particle_speed_max=1.5,
        sprite_animated=0, sprite_looped=0, sprite_frame_random_offset=3, sprite_frame_duration=0,
        texture_width=1024.0, texture_height=256.0, blend_mode=abi.BLEND_ADDITIVE, **kw)
    return p, (3, 256, 256)


def smoke(**kw):
    p = abi.default_params(
        particle_count_max=5000, emission_rate=20, ellipsoid=0,
        size_start_min=3.5, size_start_max=3.5, size_end_min=15.0, size_end_max=15.0,
        energy_min=4000.0, energy_max=5000.0,
        color_start=[0.5, 0.5, 0.5, 1], color_start_var=[0, 0, 0, 0], color_end=[1, 0, 0, 0], color_end_var=[0, 0, 0, 0],
        position=[0, 0, 0], position_var=[0, 0, 0], velocity=[2.5, 5, 0], velocity_var=[0, 0, 0],
        acceleration=[-10, 5, 0], acceleration_var=[2.5, 5, 0],
        rotation_per_particle_speed_min=-1.5, rotation_per_particle_speed_max=1.5,
        sprite_animated=0, sprite_looped=0, sprite_frame_random_offset=0, sprite_frame_duration=0,
        texture_width=64.0, texture_height=64.0, blend_mode=abi.BLEND_ADDITIVE, **kw)
    return p, (1, 64, 64)


def explosion(**kw):
    p = abi.default_params(
        particle_count_max=5000, emission_rate=50, ellipsoid=0,
        size_start_min=5.0, size_start_max=5.0, size_end_min=5.0, size_end_max=5.0,
        energy_min=1000.0, energy_max=1200.0,
        color_start=[1, 1, 1, 0.5], color_start_var=[0, 0, 0, 0], color_end=[1, 1, 1, 0], color_end_var=[0, 0, 0, 0],
        position=[0, 0, 0], position_var=[2, 2, 2], velocity=[0, 0, 0], velocity_var=[6, 6, 6],
        acceleration=[0, 0, 0], acceleration_var=[0, 0, 0],
        rotation_per_particle_speed_min=-0.5, rotation_per_particle_speed_max=0.5,
        sprite_animated=1, sprite_looped=0, sprite_frame_random_offset=4, sprite_frame_duration=25,
        texture_width=256.0, texture_height=256.0, blend_mode=abi.BLEND_ADDITIVE, **kw)
    return p, (16, 64, 64)


def headline_64m(capacity=64 * 1024 * 1024, **kw):
    """BASELINE.json configs[2] / SURVEY.md §8(d) C3: one emitter, ellipsoid spawn, per-particle spin,
    looped 4-frame sprite animation, energies long enough that nothing dies."""
    d = dict(
        particle_count_max=capacity, emission_rate=1000, ellipsoid=1,
        size_start_min=0.5, size_start_max=1.5, size_end_min=0.25, size_end_max=0.75,
        energy_min=1.0e6, energy_max=2.0e6,
        color_start=[1, 0.6, 0.2, 1], color_start_var=[0.2, 0.2, 0.1, 0], color_end=[0.2, 0.1, 0.8, 0], color_end_var=[0.1, 0.1, 0.1, 0],
        position=[0, 0, 0], position_var=[1, 1, 1], velocity=[0, 1, 0], velocity_var=[0.5, 0.5, 0.5],
        acceleration=[0, -0.1, 0], acceleration_var=[0.05, 0.05, 0.05],
        rotation_per_particle_speed_min=-1.5, rotation_per_particle_speed_max=1.5,
        sprite_animated=1, sprite_looped=1, sprite_frame_random_offset=4, sprite_frame_duration=50,
        texture_width=1024.0, texture_height=256.0, blend_mode=abi.BLEND_ADDITIVE)
    d.update(kw)
    return abi.default_params(**d), (4, 256, 256)


MIXED = (fire, smoke, explosion)
