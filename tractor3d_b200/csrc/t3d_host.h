// t3d_host.h -- pure host helpers of the particle path (no CUDA): the scalar logic the reference keeps in
// ParticleEmitter's host prologues and setters, and the `.particle` file format.  The public ones are
// declared in include/t3d_particles.h; these are the extra internal entry points.
#pragma once

#include <cstddef>
#include <cstdint>

#include "../../include/t3d_particles.h"

// PE.cpp:636-641: does the draw triple land inside the unit ball (i.e. does the rejection loop stop)?
bool t3d_host_ellipsoid_accept(int32_t rx, int32_t ry, int32_t rz);
// PE.cpp:537-546: rects {x, y, w, h} in texels -> {u1, v1, u2, v2}.
void t3d_host_sprite_rect_coords(uint32_t frame_count, const float* rects, float tex_w_ratio, float tex_h_ratio, float* out);
const char* t3d_host_last_error();
