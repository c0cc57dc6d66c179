// t3d_host.cpp -- pure host helpers (see t3d_host.h).  Compiled without FMA contraction so the few float
// expressions here round exactly like the reference's scalar SSE code.
//
// PE.cpp = Tractor3Dlib/src/graphics/ParticleEmitter.cpp, Properties.cpp = src/scene/Properties.cpp,
// ParticlesSample.cpp = samples/browser/src/ParticlesSample.cpp of the reference.
#include "t3d_host.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "t3d_internal.h"

static thread_local std::string g_host_error;
const char* t3d_host_last_error() { return g_host_error.c_str(); }
static int host_fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_host_error = buf;
    return code;
}

static inline float u11(int32_t r) { return (2.0f * ((float)r / 2147483648.0f)) - 1.0f; } // pch.h:151

bool t3d_host_ellipsoid_accept(int32_t rx, int32_t ry, int32_t rz)
{
    const float x = u11(rx), y = u11(ry), z = u11(rz);
    return !(sqrtf(x * x + y * y + z * z) > 1.0f); // Vector3::length() > 1.0f keeps looping
}

void t3d_host_sprite_rect_coords(uint32_t frame_count, const float* rects, float wr, float hr, float* out)
{
    for (uint32_t i = 0; i < frame_count; ++i)
    {
        const float* r = rects + 4 * i;
        float* c = out + 4 * i;
        c[0] = wr * r[0];
        c[1] = 1.0f - hr * r[1];
        c[2] = c[0] + wr * r[2];
        c[3] = c[1] - hr * r[3];
    }
}

extern "C" {

void t3d_emitter_params_default(t3d_emitter_params* p)
{
    // PE.h:824-877 defaults; capacity/rate as create(Properties*) falls back to (PE.cpp:133-143).
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->particle_count_max = 100;
    p->emission_rate = 10;
    p->size_start_min = p->size_start_max = p->size_end_min = p->size_end_max = 1.0f;
    p->energy_min = p->energy_max = 1000.0f;
    for (int k = 0; k < 4; ++k) p->color_end[k] = 1.0f;
    for (int k = 0; k < 3; ++k) p->velocity_var[k] = 1.0f;
    p->texture_width = p->texture_height = 64.0f;
    p->blend_mode = T3D_BLEND_ALPHA;
}

int32_t t3d_device_rng_draw(uint64_t seed, uint64_t serial, uint32_t j) { return t3d::device_rng_draw(seed, serial, j); }

int t3d_host_fuse_policy(uint64_t particles, uint64_t spawned)
{
    // Measured on B200 (profiles/r2_churn.txt): since a moved particle leaves with its destination group's vector stores
    // and its quad with the warp's bulk store, the fused frame beats the two launches at every churn level -- by 8 % with
    // nothing replaced per frame, 9 % at 0.34 %, 8 % at 1.3-1.7 %, 4 % at 7 % -- so update() always waits for the draw()
    // behind it.  (Round 1's kernel paid a scattered 144-byte quad per moved particle and lost beyond 0.25 %.)
    (void)particles;
    (void)spawned;
    return 1;
}

int t3d_host_rate_cap(double* running_time, float elapsed_time, float* elapsed_ms)
{
    // PE.cpp:720-725: PARTICLE_UPDATE_RATE_MAX = 8 ms.
    *running_time += elapsed_time;
    if (*running_time < 8) return 0;
    *elapsed_ms = (float)*running_time;
    *running_time = 0;
    return 1;
}

uint32_t t3d_host_emission_count(float* emit_time, float time_per_emission, float elapsed_ms)
{
    // PE.cpp:732-743.  For rates above 1000/s (int)time_per_emission is 0 and emit_time is never reduced.
    *emit_time += elapsed_ms;
    const unsigned int emit_count = (unsigned int)(*emit_time / time_per_emission);
    if (emit_count)
    {
        if ((int)time_per_emission > 0) *emit_time = (float)fmod((double)*emit_time, (double)time_per_emission);
    }
    return emit_count;
}

// ui Sprite::draw, src/ui/Sprite.cpp:495-574, up to its SpriteBatch::draw call, and that call's own first lines (SB.cpp:202-216).
void t3d_host_ui_sprite_record(const t3d_ui_sprite* s, uint32_t texture_width, uint32_t texture_height, t3d_sprite2d* out)
{
    if (!s || !out) return;
    float px = 0.0f, py = 0.0f, pz = 0.0f;                           // :498 Vector3::zero()
    px -= s->camera_translation[0];                                  // :514-515
    py -= s->camera_translation[1];
    px += s->node_translation[0];                                    // :520-523
    py += s->node_translation[1];
    pz += s->node_translation[2];
    const uint32_t off = s->offset;
    if ((off & 0x02u) == 0x02u) px = (float)((double)px - (double)s->width * 0.5); // :527 (`_width * 0.5`: a double product)
    if ((off & 0x04u) == 0x04u) px -= s->width;                      // :528
    if ((off & 0x20u) == 0x20u) py -= s->height * 0.5f;              // :529
    if ((off & 0x10u) == 0x10u) py -= s->height;                     // :530
    if ((off & 0x80u) == 0x80u)                                      // :531-535
    {
        px -= s->width * s->anchor[0];
        py -= s->height * s->anchor[1];
    }
    float angle = 0.0f;                                              // :538
    float sx = s->width, sy = s->height;                             // :539
    const float qx = s->node_rotation[0], qy = s->node_rotation[1], qz = s->node_rotation[2];
    if (qx != 0.0f || qy != 0.0f || qz != 0.0f)                      // :543-544: Quaternion::toAxisAngle(nullptr)
    {
        float w = s->node_rotation[3];
        float n = qx * qx + qy * qy + qz * qz + w * w;               // src/math/Quaternion.cpp:183-196 (normalize)
        if (n != 1.0f)
        {
            n = sqrtf(n);
            if (!(n < 0.000001f)) w *= 1.0f / n;
        }
        angle = (float)(2.0 * acos((double)w));                      // Quaternion.cpp:277: `2.0f * acos(q.w)`, the double acos
    }
    if (s->node_scale[0] != 1.0f) sx *= s->node_scale[0];            // :547-548
    if (s->node_scale[1] != 1.0f) sy *= s->node_scale[1];
    if ((s->flip & 2u) == 2u)                                        // :551-555 FLIP_HORIZONTAL
    {
        px += sx;
        sx = -sx;
    }
    if ((s->flip & 1u) == 1u)                                        // :557-561 FLIP_VERTICAL
    {
        py += sy;
        sy = -sy;
    }
    memset(out, 0, sizeof(*out));
    const float wr = 1.0f / (float)texture_width, hr = 1.0f / (float)texture_height; // SB.cpp:155-156
    out->x = px; out->y = py; out->z = pz;
    out->width = sx; out->height = sy;
    out->u1 = wr * s->frame[0];                                      // SB.cpp:208-211
    out->v1 = 1.0f - hr * s->frame[1];
    out->u2 = out->u1 + wr * s->frame[2];
    out->v2 = out->v1 - hr * s->frame[3];
    out->color[0] = s->color[0]; out->color[1] = s->color[1]; out->color[2] = s->color[2];
    out->color[3] = s->color[3] * s->opacity;                        // :568
    out->rotation_point[0] = s->anchor[0];                           // :569
    out->rotation_point[1] = s->anchor[1];
    out->rotation_angle = angle;
    out->flags = T3D_SPRITE_ROTATED;
}

// Font::measureText(text, size, width, height), Font.cpp:675-731; the width of a line is Font::getTokenWidth, :1624-1657.
void t3d_host_font_measure_text(const t3d_glyph* glyphs, uint32_t glyph_count, uint32_t font_size, float character_spacing, const char* text,
                                size_t length, uint32_t size, uint32_t* width, uint32_t* height)
{
    uint32_t w = 0, h = 0;
    if (size == 0) size = font_size;                                   // :684-687
    size_t end = 0;                                                    // the walk is over text.c_str(): it ends at the first NUL
    while (end < length && text[end] != 0) ++end;
    if (length != 0 && glyphs && glyph_count && font_size)             // :699-705: an empty string measures 0 x 0
    {
        const float scale = (float)size / (float)font_size;            // :707
        const int spacing = (int)((float)size * character_spacing);    // :1630
        h = size;                                                      // :711
        size_t i = 0;
        while (i < end)                                                // :714
        {
            while (i < end && text[i] == '\n')                         // :716-720
            {
                h += size;
                ++i;
            }
            uint32_t line = 0;                                         // :722-723, getTokenWidth over the bytes up to the next '\n'
            for (; i < end && text[i] != '\n'; ++i)
            {
                const signed char c = (signed char)text[i];            // `char`, signed on the reference's platforms
                if (c == ' ') line += glyphs[0].advance;               // :1638-1640
                else if (c == '\t') line += glyphs[0].advance * 4u;    // :1641-1643
                else
                {
                    const int index = c - 32;                          // :1645
                    if (index >= 0 && index < (int)glyph_count)
                        // :1649 `tokenWidth += floor(...)`: the floor in float, the sum in double, back to unsigned
                        line = (uint32_t)(long long)((double)line + (double)floorf((float)glyphs[index].advance * scale + (float)spacing));
                }
            }
            if (line > w) w = line;                                    // :724-727
        }
    }
    if (width) *width = w;
    if (height) *height = h;
}

uint32_t t3d_host_sprite_frame_coords(uint32_t frame_count, int width, int height, float tex_w, float tex_h, float* tex_coords)
{
    // PE.cpp:550-582: frames walk the sheet left-to-right, top-to-bottom; rects the walk never reaches stay
    // zero (Rectangle's default), exactly like the reference's `new Rectangle[frameCount]`.
    if (!frame_count || !width || !height || !tex_coords) return 0;
    std::vector<float> rects((size_t)frame_count * 4, 0.0f);
    const unsigned int cols = (unsigned int)(tex_w / (float)width);
    const unsigned int rows = (unsigned int)(tex_h / (float)height);
    unsigned int n = 0;
    for (unsigned int i = 0; i < rows && n != frame_count; ++i)
    {
        const int y = (int)i * height;
        for (unsigned int j = 0; j < cols; ++j)
        {
            const int x = (int)j * width;
            float* r = &rects[4 * ((size_t)i * cols + j)];
            r[0] = (float)x;
            r[1] = (float)y;
            r[2] = (float)width;
            r[3] = (float)height;
            if (++n == frame_count) break;
        }
    }
    const float wr = 1.0f / (float)(unsigned int)tex_w, hr = 1.0f / (float)(unsigned int)tex_h; // PE.cpp:246-247
    t3d_host_sprite_rect_coords(frame_count, rects.data(), wr, hr, tex_coords);
    return n;
}

size_t t3d_host_particle_draws(const int32_t* draws, size_t n, int ellipsoid, uint32_t frame_random_offset)
{
    // PE.cpp:312-364: 8 colour + 6 scalar draws, the position (3 per rejection-loop pass), 9 for
    // velocity/acceleration/axis, one more when a random start frame is drawn.
    size_t pos = 14;
    if (ellipsoid)
    {
        while (true)
        {
            if (pos + 3 > n) return 0;
            const bool ok = t3d_host_ellipsoid_accept(draws[pos], draws[pos + 1], draws[pos + 2]);
            pos += 3;
            if (ok) break;
        }
    }
    else
        pos += 3;
    pos += 9 + (frame_random_offset > 0 ? 1 : 0);
    return pos <= n ? pos : 0;
}

} // extern "C"

// --------------------------------------------------------------------------------------------------
// `.particle` files: the Properties text format restricted to what ParticleEmitter::create reads
// (PE.cpp:92-184): nested namespaces `name [id] { ... }` and `key = value` lines, // and /* */ comments.
// --------------------------------------------------------------------------------------------------
namespace
{
struct Ns
{
    std::string name, id;
    std::vector<std::pair<std::string, std::string>> props;
    std::vector<std::unique_ptr<Ns>> children;
    const std::string* find(const std::string& key) const
    {
        for (auto& kv : props)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Ns* child(const std::string& n) const
    {
        for (auto& c : children)
            if (c->name == n) return c.get();
        return nullptr;
    }
};

std::string trim(const std::string& s)
{
    const auto a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    const auto b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}

bool parse_file(const std::string& path, Ns* root, std::string* err)
{
    std::ifstream f(path);
    if (!f)
    {
        *err = "cannot open '" + path + "'";
        return false;
    }
    std::vector<Ns*> stack{ root };
    Ns* awaiting_brace = nullptr; // a namespace header was read; its '{' is on a later line
    bool in_comment = false;
    std::string raw;
    while (std::getline(f, raw))
    {
        std::string line = trim(raw);
        if (line.empty()) continue;
        if (in_comment)
        {
            if (line.size() >= 2 && (line.substr(0, 2) == "*/" || line.substr(line.size() - 2) == "*/")) in_comment = false;
            continue;
        }
        if (line.size() >= 2 && line.substr(0, 2) == "/*")
        {
            if (!(line.size() >= 4 && line.substr(line.size() - 2) == "*/")) in_comment = true;
            continue;
        }
        if (line.size() >= 2 && line.substr(0, 2) == "//") continue;
        const auto eq = line.find('=');
        if (eq != std::string::npos)
        {
            stack.back()->props.emplace_back(trim(line.substr(0, eq)), trim(line.substr(eq + 1)));
            continue;
        }
        if (line[0] == '{')
        {
            if (!awaiting_brace)
            {
                *err = "unexpected '{'";
                return false;
            }
            stack.push_back(awaiting_brace);
            awaiting_brace = nullptr;
            continue;
        }
        if (line[0] == '}')
        {
            if (stack.size() <= 1)
            {
                *err = "unbalanced '}'";
                return false;
            }
            stack.pop_back();
            continue;
        }
        // namespace header: name [id] [: parent] [{] [}]
        const bool open = line.find('{') != std::string::npos, close = line.find('}') != std::string::npos;
        std::string head = line.substr(0, line.find_first_of("{:"));
        std::istringstream hs(head);
        auto ns = std::make_unique<Ns>();
        hs >> ns->name >> ns->id;
        Ns* np = ns.get();
        stack.back()->children.push_back(std::move(ns));
        if (open && !close)
            stack.push_back(np);
        else if (!open)
            awaiting_brace = np;
    }
    return true;
}

// Typed getters with Properties.cpp's conversions: getBool :940-944, getInt :947-962, getFloat :965-978,
// getLong :981-994, getVector3/4 via sscanf :1214-1251.
bool get_bool(const Ns* n, const char* k)
{
    const std::string* v = n->find(k);
    return v && !v->empty() && *v == "true";
}
int get_int(const Ns* n, const char* k)
{
    const std::string* v = n->find(k);
    if (!v || v->empty()) return 0;
    try { return std::stoi(*v); } catch (...) { return 0; }
}
long get_long(const Ns* n, const char* k)
{
    const std::string* v = n->find(k);
    if (!v) return 0;
    try { return std::stol(*v); } catch (...) { return 0; }
}
float get_float(const Ns* n, const char* k)
{
    const std::string* v = n->find(k);
    if (!v) return 0.0f;
    try { return std::stof(*v); } catch (...) { return 0.0f; }
}
void get_vec(const Ns* n, const char* k, float* out, int dim)
{
    // The caller's vector keeps its default (zero) when the key is missing; a malformed value zeroes it.
    for (int i = 0; i < dim; ++i) out[i] = 0.0f;
    const std::string* v = n->find(k);
    if (!v || v->empty()) return;
    float t[4] = { 0, 0, 0, 0 };
    const int got = dim == 3 ? sscanf(v->c_str(), "%f,%f,%f", &t[0], &t[1], &t[2])
                             : sscanf(v->c_str(), "%f,%f,%f,%f", &t[0], &t[1], &t[2], &t[3]);
    if (got == dim)
        for (int i = 0; i < dim; ++i) out[i] = t[i];
}
int blend_from_string(const std::string& s)
{
    // PE.cpp:682-710
    if (s == "BLEND_NONE" || s == "NONE" || s == "BLEND_OPAQUE" || s == "OPAQUE") return T3D_BLEND_NONE;
    if (s == "BLEND_ALPHA" || s == "ALPHA" || s == "BLEND_TRANSPARENT" || s == "TRANSPARENT") return T3D_BLEND_ALPHA;
    if (s == "BLEND_ADDITIVE" || s == "ADDITIVE") return T3D_BLEND_ADDITIVE;
    if (s == "BLEND_MULTIPLIED" || s == "MULTIPLIED") return T3D_BLEND_MULTIPLIED;
    return T3D_BLEND_ALPHA;
}
bool file_exists(const std::string& p)
{
    FILE* f = fopen(p.c_str(), "rb");
    if (f) fclose(f);
    return f != nullptr;
}
bool png_size(const std::string& p, unsigned* w, unsigned* h)
{
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) return false;
    unsigned char hdr[24];
    const size_t got = fread(hdr, 1, sizeof(hdr), f);
    fclose(f);
    if (got != sizeof(hdr) || memcmp(hdr + 1, "PNG", 3) != 0) return false;
    *w = (hdr[16] << 24) | (hdr[17] << 16) | (hdr[18] << 8) | hdr[19];
    *h = (hdr[20] << 24) | (hdr[21] << 16) | (hdr[22] << 8) | hdr[23];
    return *w && *h;
}
} // namespace

extern "C" int t3d_host_png_size(const char* path, uint32_t* width, uint32_t* height)
{
    unsigned w = 0, h = 0;
    if (!path || !png_size(path, &w, &h)) return host_fail(T3D_ERR_IO, "Failed to create texture for particle emitter: '%s'", path ? path : "");
    if (width) *width = w;
    if (height) *height = h;
    return T3D_OK;
}

// A parsed (or hand-built) namespace tree handed across the ABI.  The root owns the whole tree.
struct t3d_properties
{
    Ns* ns{ nullptr };          // this namespace
    std::string dir;            // directory of the file it came from ("" for hand-built trees): relative paths resolve here
    std::unique_ptr<Ns> owned;  // set on roots
    std::vector<std::unique_ptr<t3d_properties>> views; // child handles given out, kept alive with the root
    t3d_properties* root{ nullptr };
};

// PE.cpp:92-211 on a `particle` namespace: its first child must be `sprite`; conversions as Properties.cpp's getters.
static int params_from_namespace(const Ns* particle, const std::string& dir, t3d_emitter_params* params, uint32_t* frame_count,
                                 int* sprite_width, int* sprite_height, char* texture_path, size_t texture_path_len)
{
    if (!particle || particle->name != "particle")
        return host_fail(T3D_ERR_IO, "Properties object must be non-null and have namespace equal to 'particle'.");
    // PE.cpp:101-106: the first child namespace must be `sprite`.
    const Ns* sprite = particle->children.empty() ? nullptr : particle->children.front().get();
    if (!sprite || sprite->name != "sprite")
        return host_fail(T3D_ERR_IO, "Failed to load particle emitter: required namespace 'sprite' is missing.");
    // PE.cpp:110-115 + Properties::getPath (Properties.cpp:1077-1108): as given, else relative to the file.
    const std::string* path = sprite->find("path");
    if (!path || path->empty())
        return host_fail(T3D_ERR_IO, "Failed to load particle emitter: required image file path ('path') is missing.");
    std::string tex = *path;
    if (!file_exists(tex))
    {
        if (!dir.empty() && file_exists(dir + tex))
            tex = dir + tex;
        else
            return host_fail(T3D_ERR_IO, "Failed to load particle emitter: required image file path ('path') is missing.");
    }
    unsigned tw = 0, th = 0;
    if (!png_size(tex, &tw, &th)) return host_fail(T3D_ERR_IO, "Failed to create texture for particle emitter.");

    t3d_emitter_params p;
    t3d_emitter_params_default(&p);
    const std::string* bm = sprite->find("blendMode");
    if (!bm || bm->empty()) bm = sprite->find("blending"); // PE.cpp:117-122
    p.blend_mode = blend_from_string(bm ? *bm : std::string());
    const int sw = get_int(sprite, "width"), sh = get_int(sprite, "height");
    p.sprite_animated = get_bool(sprite, "animated");
    p.sprite_looped = get_bool(sprite, "looped");
    const int fc = get_int(sprite, "frameCount");
    p.sprite_frame_random_offset = (uint32_t)std::min(get_int(sprite, "frameRandomOffset"), fc); // PE.cpp:128
    p.sprite_frame_duration = (long)get_float(sprite, "frameDuration"); // float -> setSpriteFrameDuration(long), PE.cpp:129, 206
    p.particle_count_max = (unsigned int)get_int(particle, "particleCountMax");
    if (p.particle_count_max == 0) p.particle_count_max = 100; // PE.cpp:133-137
    p.emission_rate = (unsigned int)get_int(particle, "emissionRate");
    if (p.emission_rate == 0) p.emission_rate = 10; // PE.cpp:139-143
    p.ellipsoid = get_bool(particle, "ellipsoid");
    p.size_start_min = get_float(particle, "sizeStartMin");
    p.size_start_max = get_float(particle, "sizeStartMax");
    p.size_end_min = get_float(particle, "sizeEndMin");
    p.size_end_max = get_float(particle, "sizeEndMax");
    p.energy_min = (float)get_long(particle, "energyMin"); // long -> setEnergy(long, long) -> float members, PE.cpp:150-151, 381-385
    p.energy_max = (float)get_long(particle, "energyMax");
    struct { const char* key; float* dst; int dim; } vecs[] = {
        { "colorStart", p.color_start, 4 }, { "colorStartVar", p.color_start_var, 4 }, { "colorEnd", p.color_end, 4 },
        { "colorEndVar", p.color_end_var, 4 }, { "position", p.position, 3 }, { "positionVar", p.position_var, 3 },
        { "velocity", p.velocity, 3 }, { "velocityVar", p.velocity_var, 3 }, { "acceleration", p.acceleration, 3 },
        { "accelerationVar", p.acceleration_var, 3 }, { "rotationAxis", p.rotation_axis, 3 }, { "rotationAxisVar", p.rotation_axis_var, 3 },
    };
    for (auto& v : vecs) get_vec(particle, v.key, v.dst, v.dim);
    p.rotation_per_particle_speed_min = get_float(particle, "rotationPerParticleSpeedMin");
    p.rotation_per_particle_speed_max = get_float(particle, "rotationPerParticleSpeedMax");
    p.rotation_speed_min = get_float(particle, "rotationSpeedMin");
    p.rotation_speed_max = get_float(particle, "rotationSpeedMax");
    p.orbit_position = get_bool(particle, "orbitPosition");
    p.orbit_velocity = get_bool(particle, "orbitVelocity");
    p.orbit_acceleration = get_bool(particle, "orbitAcceleration");
    p.texture_width = (float)tw;
    p.texture_height = (float)th;
    *params = p;
    if (frame_count) *frame_count = (uint32_t)fc;
    if (sprite_width) *sprite_width = sw;
    if (sprite_height) *sprite_height = sh;
    if (texture_path && texture_path_len) snprintf(texture_path, texture_path_len, "%s", tex.c_str());
    return T3D_OK;
}

static std::string directory_of(const char* url)
{
    std::string dir(url);
    const auto slash = dir.find_last_of('/');
    return slash == std::string::npos ? "" : dir.substr(0, slash + 1);
}

extern "C" int t3d_particle_file_parse(const char* url, t3d_emitter_params* params, uint32_t* frame_count, int* sprite_width,
                                       int* sprite_height, char* texture_path, size_t texture_path_len)
{
    if (!url || !params) return host_fail(T3D_ERR_INVALID, "t3d_particle_file_parse: null argument");
    Ns root;
    std::string err;
    if (!parse_file(url, &root, &err)) return host_fail(T3D_ERR_IO, "Failed to create particle emitter from file: %s", err.c_str());
    // PE.cpp:85-86, 94: the (first) top-level namespace must be `particle`.
    const Ns* particle = root.children.empty() ? nullptr : root.children.front().get();
    return params_from_namespace(particle, directory_of(url), params, frame_count, sprite_width, sprite_height, texture_path, texture_path_len);
}

// ---- t3d_properties: the slice of tractor::Properties (src/scene/Properties.cpp) the particle path reads ----
static t3d_properties* view_of(t3d_properties* root, Ns* ns)
{
    for (auto& v : root->views)
        if (v->ns == ns) return v.get();
    auto v = std::make_unique<t3d_properties>();
    v->ns = ns;
    v->dir = root->dir;
    v->root = root;
    root->views.push_back(std::move(v));
    return root->views.back().get();
}

extern "C" {

int t3d_properties_load(const char* url, t3d_properties** out)
{
    if (!url || !out) return host_fail(T3D_ERR_INVALID, "t3d_properties_load: null argument");
    *out = nullptr;
    auto root = std::make_unique<t3d_properties>();
    root->owned = std::make_unique<Ns>();
    std::string err;
    if (!parse_file(url, root->owned.get(), &err)) return host_fail(T3D_ERR_IO, "Failed to load properties from '%s': %s", url, err.c_str());
    root->ns = root->owned.get();
    root->dir = directory_of(url);
    root->root = root.get();
    *out = root.release();
    return T3D_OK;
}

int t3d_properties_new(const char* name_space, const char* id, t3d_properties** out)
{
    if (!out) return host_fail(T3D_ERR_INVALID, "t3d_properties_new: null argument");
    auto root = std::make_unique<t3d_properties>();
    root->owned = std::make_unique<Ns>();
    root->owned->name = name_space ? name_space : "";
    root->owned->id = id ? id : "";
    root->ns = root->owned.get();
    root->root = root.get();
    *out = root.release();
    return T3D_OK;
}

void t3d_properties_free(t3d_properties* root)
{
    if (root && root->root == root) delete root;
}

int t3d_properties_set(t3d_properties* p, const char* key, const char* value)
{
    if (!p || !key) return host_fail(T3D_ERR_INVALID, "t3d_properties_set: null argument");
    for (auto& kv : p->ns->props)
        if (kv.first == key)
        {
            kv.second = value ? value : "";
            return T3D_OK;
        }
    p->ns->props.emplace_back(key, value ? value : "");
    return T3D_OK;
}

int t3d_properties_add_namespace(t3d_properties* p, const char* name_space, const char* id, t3d_properties** child)
{
    if (!p || !name_space) return host_fail(T3D_ERR_INVALID, "t3d_properties_add_namespace: null argument");
    auto ns = std::make_unique<Ns>();
    ns->name = name_space;
    ns->id = id ? id : "";
    Ns* raw = ns.get();
    p->ns->children.push_back(std::move(ns));
    if (child) *child = view_of(p->root, raw);
    return T3D_OK;
}

const char* t3d_properties_namespace(const t3d_properties* p) { return p ? p->ns->name.c_str() : ""; }
const char* t3d_properties_id(const t3d_properties* p) { return p ? p->ns->id.c_str() : ""; }
size_t t3d_properties_namespace_count(const t3d_properties* p) { return p ? p->ns->children.size() : 0; }
t3d_properties* t3d_properties_get_namespace(t3d_properties* p, size_t index)
{
    if (!p || index >= p->ns->children.size()) return nullptr;
    return view_of(p->root, p->ns->children[index].get());
}
size_t t3d_properties_count(const t3d_properties* p) { return p ? p->ns->props.size() : 0; }
const char* t3d_properties_key(const t3d_properties* p, size_t index)
{
    return (p && index < p->ns->props.size()) ? p->ns->props[index].first.c_str() : nullptr;
}
const char* t3d_properties_value(const t3d_properties* p, size_t index)
{
    return (p && index < p->ns->props.size()) ? p->ns->props[index].second.c_str() : nullptr;
}
const char* t3d_properties_get(const t3d_properties* p, const char* key)
{
    if (!p || !key) return nullptr;
    const std::string* v = p->ns->find(key);
    return v ? v->c_str() : nullptr;
}

int t3d_properties_get_bool(const t3d_properties* p, const char* key, int default_value)
{
    if (!p || !key) return default_value;
    const std::string* v = p->ns->find(key);
    return (v && !v->empty()) ? (*v == "true") : default_value; // Properties.cpp:940-944
}
int t3d_properties_get_int(const t3d_properties* p, const char* key) { return (p && key) ? get_int(p->ns, key) : 0; }
long long t3d_properties_get_long(const t3d_properties* p, const char* key) { return (p && key) ? get_long(p->ns, key) : 0; }
float t3d_properties_get_float(const t3d_properties* p, const char* key) { return (p && key) ? get_float(p->ns, key) : 0.0f; }
int t3d_properties_get_vector(const t3d_properties* p, const char* key, float* out, int dim)
{
    if (!p || !key || !out || (dim != 3 && dim != 4)) return 0;
    const std::string* v = p->ns->find(key);
    get_vec(p->ns, key, out, dim);
    if (!v || v->empty()) return 0;
    float t[4];
    return (dim == 3 ? sscanf(v->c_str(), "%f,%f,%f", &t[0], &t[1], &t[2]) : sscanf(v->c_str(), "%f,%f,%f,%f", &t[0], &t[1], &t[2], &t[3])) == dim;
}
int t3d_properties_get_path(const t3d_properties* p, const char* key, char* out, size_t out_len)
{
    if (!p || !key || !out || !out_len) return host_fail(T3D_ERR_INVALID, "t3d_properties_get_path: null argument");
    const std::string* v = p->ns->find(key);
    if (!v || v->empty()) return host_fail(T3D_ERR_IO, "no property '%s'", key);
    std::string path = *v;
    if (!file_exists(path))
    {
        if (p->dir.empty() || !file_exists(p->dir + path)) return host_fail(T3D_ERR_IO, "'%s' does not name an existing file", v->c_str());
        path = p->dir + path;
    }
    snprintf(out, out_len, "%s", path.c_str());
    return T3D_OK;
}

int t3d_particle_properties_parse(const t3d_properties* particle, t3d_emitter_params* params, uint32_t* frame_count, int* sprite_width,
                                  int* sprite_height, char* texture_path, size_t texture_path_len)
{
    if (!particle || !params) return host_fail(T3D_ERR_INVALID, "t3d_particle_properties_parse: null argument");
    return params_from_namespace(particle->ns, particle->dir, params, frame_count, sprite_width, sprite_height, texture_path, texture_path_len);
}

} // extern "C"

extern "C" int t3d_particle_file_save(const char* url, const char* name, const t3d_emitter_params* p, uint32_t frame_count,
                                      int sprite_width, int sprite_height, const char* texture_path)
{
    // ParticlesSample.cpp:340-424 (saveFile): same keys, same order; like the sample it omits the
    // rotationSpeed* / rotationAxis* keys.  The trailing `editor` block of the sample is UI state and is not written.
    if (!url || !p) return host_fail(T3D_ERR_INVALID, "t3d_particle_file_save: null argument");
    static const char* kBlend[] = { "NONE", "ALPHA", "ADDITIVE", "MULTIPLIED" };
    auto B = [](int v) { return v ? "true" : "false"; };
    std::ostringstream s; // default ostream float formatting, as the sample
    auto V3 = [&](const float* v) { std::ostringstream t; t << v[0] << ", " << v[1] << ", " << v[2]; return t.str(); };
    auto V4 = [&](const float* v) { std::ostringstream t; t << v[0] << ", " << v[1] << ", " << v[2] << ", " << v[3]; return t.str(); };
    s << "particle " << (name ? name : "") << "\n{\n    sprite\n    {\n"
      << "        path = " << (texture_path ? texture_path : "") << "\n"
      << "        width = " << sprite_width << "\n"
      << "        height = " << sprite_height << "\n"
      << "        blendMode = " << kBlend[(p->blend_mode >= 0 && p->blend_mode < 4) ? p->blend_mode : 1] << "\n"
      << "        animated = " << B(p->sprite_animated) << "\n"
      << "        looped = " << B(p->sprite_looped) << "\n"
      << "        frameCount = " << frame_count << "\n"
      << "        frameRandomOffset = " << (int)p->sprite_frame_random_offset << "\n"
      << "        frameDuration = " << (long)p->sprite_frame_duration << "\n"
      << "    }\n\n"
      << "    particleCountMax = " << p->particle_count_max << "\n"
      << "    emissionRate = " << p->emission_rate << "\n"
      << "    ellipsoid = " << B(p->ellipsoid) << "\n"
      << "    orbitPosition = " << B(p->orbit_position) << "\n"
      << "    orbitVelocity = " << B(p->orbit_velocity) << "\n"
      << "    orbitAcceleration = " << B(p->orbit_acceleration) << "\n"
      << "    sizeStartMin = " << p->size_start_min << "\n"
      << "    sizeStartMax = " << p->size_start_max << "\n"
      << "    sizeEndMin = " << p->size_end_min << "\n"
      << "    sizeEndMax = " << p->size_end_max << "\n"
      << "    energyMin = " << (long)p->energy_min << "\n" // getEnergyMin() returns long, PE.h:427
      << "    energyMax = " << (long)p->energy_max << "\n"
      << "    colorStart = " << V4(p->color_start) << "\n"
      << "    colorStartVar = " << V4(p->color_start_var) << "\n"
      << "    colorEnd = " << V4(p->color_end) << "\n"
      << "    colorEndVar = " << V4(p->color_end_var) << "\n"
      << "    position = " << V3(p->position) << "\n"
      << "    positionVar = " << V3(p->position_var) << "\n"
      << "    velocity = " << V3(p->velocity) << "\n"
      << "    velocityVar = " << V3(p->velocity_var) << "\n"
      << "    acceleration = " << V3(p->acceleration) << "\n"
      << "    accelerationVar = " << V3(p->acceleration_var) << "\n"
      << "    rotationPerParticleSpeedMin = " << p->rotation_per_particle_speed_min << "\n"
      << "    rotationPerParticleSpeedMax = " << p->rotation_per_particle_speed_max << "\n"
      << "}\n";
    std::ofstream f(url);
    if (!f) return host_fail(T3D_ERR_IO, "cannot write '%s'", url);
    f << s.str();
    return f.good() ? T3D_OK : host_fail(T3D_ERR_IO, "write to '%s' failed", url);
}
