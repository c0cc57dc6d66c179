// multi_ctx_driver.cpp -- what a C++ engine with several GPUs does: ONE process, ONE thread, one t3d_ctx per GPU, the
// emitters split by ownership (emitter e -> context e mod N), everything through the C ABI of include/t3d_particles.h.
//
//   multi_ctx_driver <params.bin> <n_contexts> <same|spread> <dt_ms> <frames> <out.bin> [private]
//
// params.bin: K records of { t3d_emitter_params, uint32 frame_count, int32 sprite_w, int32 sprite_h, uint64 seed }.
// Runs the frames twice -- all emitters on one context, then split over n_contexts (all on device 0 with `same`,
// context k on device k with `spread`) -- with the same call order, and requires the two runs to agree bit for bit:
// the rate-cap accumulator (PE.cpp:720) and the latched billboard rotation (SB.cpp:339) are process-wide, so the split
// must not be observable.  The split run's states are written to out.bin for the oracle comparison in
// tests/test_gpu_multi_context.py.  With `private` every context keeps its own statics and the driver reports whether
// the split became observable (it does as soon as an update step is under 8 ms).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/t3d_particles.h"

#define CHECK(call)                                                                                   \
    do                                                                                                \
    {                                                                                                 \
        if ((call) != T3D_OK)                                                                         \
        {                                                                                             \
            std::fprintf(stderr, "%s:%d %s failed: %s\n", __FILE__, __LINE__, #call, t3d_last_error()); \
            std::exit(10);                                                                            \
        }                                                                                             \
    } while (0)

struct Record
{
    t3d_emitter_params p;
    uint32_t frame_count;
    int32_t sprite_w, sprite_h;
    int32_t pad;
    uint64_t seed;
};
static_assert(sizeof(Record) == sizeof(t3d_emitter_params) + 24, "the layout tests/test_gpu_multi_context.py writes");

struct Result
{
    std::vector<uint32_t> counts;
    std::vector<std::vector<uint32_t>> states; // T3D_NUM_COLUMNS x count, per emitter
    std::vector<std::vector<float>> verts;
};

static Result run(const std::vector<Record>& recs, int n_ctx, bool spread, float dt, int frames, bool private_statics)
{
    std::vector<t3d_ctx*> ctxs(n_ctx);
    for (int k = 0; k < n_ctx; ++k) CHECK(t3d_ctx_create(spread ? k : 0, nullptr, &ctxs[k]));
    CHECK(t3d_ctx_reset_statics(ctxs[0])); // "a fresh process"
    if (private_statics)
        for (t3d_ctx* c : ctxs) CHECK(t3d_ctx_set_private_statics(c, 1));
    const size_t K = recs.size();
    std::vector<t3d_emitter*> es(K);
    float cam[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.2f, 2.5f, 13.0f, 1 };
    for (size_t e = 0; e < K; ++e)
    {
        CHECK(t3d_emitter_create(ctxs[e % n_ctx], &recs[e].p, &es[e]));
        CHECK(t3d_emitter_set_sprite_frame_coords(es[e], recs[e].frame_count, recs[e].sprite_w, recs[e].sprite_h));
        CHECK(t3d_emitter_set_rng(es[e], T3D_RNG_DEVICE, recs[e].seed));
        CHECK(t3d_emitter_set_camera_world(es[e], cam));
        CHECK(t3d_emitter_start(es[e]));
    }
    for (int f = 0; f < frames; ++f)
    {
        // the game loop: every emitter's node moves, update() on each, then draw() on each (ParticlesSample.cpp:897, 916)
        for (size_t e = 0; e < K; ++e)
        {
            float w[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.01f * (float)f * (float)(e + 1), 0.5f * (float)e, 0, 1 };
            CHECK(t3d_emitter_set_node_world(es[e], w));
        }
        for (size_t e = 0; e < K; ++e) CHECK(t3d_emitter_update(es[e], dt));
        for (size_t e = 0; e < K; ++e)
        {
            uint32_t calls = 0;
            CHECK(t3d_emitter_draw(es[e], &calls));
        }
    }
    Result r;
    for (size_t e = 0; e < K; ++e)
    {
        uint32_t n = 0;
        CHECK(t3d_emitter_get_count(es[e], &n));
        std::vector<uint32_t> st((size_t)T3D_NUM_COLUMNS * (n ? n : 1));
        uint32_t got = 0;
        CHECK(t3d_emitter_download_state(es[e], st.data(), n ? n : 1, &got));
        if (got != n) std::exit(11);
        st.resize((size_t)T3D_NUM_COLUMNS * n);
        std::vector<float> v((size_t)T3D_FLOATS_PER_QUAD * (n ? n : 1));
        uint32_t q = 0;
        CHECK(t3d_emitter_download_vertices(es[e], v.data(), n, &q));
        v.resize((size_t)T3D_FLOATS_PER_QUAD * q);
        r.counts.push_back(n);
        r.states.push_back(std::move(st));
        r.verts.push_back(std::move(v));
    }
    for (t3d_emitter* e : es) CHECK(t3d_emitter_destroy(e));
    for (t3d_ctx* c : ctxs) CHECK(t3d_ctx_destroy(c));
    return r;
}

int main(int argc, char** argv)
{
    if (argc < 7)
    {
        std::fprintf(stderr, "usage: %s params.bin n_contexts same|spread dt_ms frames out.bin [private]\n", argv[0]);
        return 2;
    }
    std::vector<Record> recs;
    {
        FILE* f = std::fopen(argv[1], "rb");
        if (!f) return 3;
        Record r;
        while (std::fread(&r, sizeof(r), 1, f) == 1) recs.push_back(r);
        std::fclose(f);
    }
    const int n_ctx = std::atoi(argv[2]);
    const bool spread = std::string(argv[3]) == "spread";
    const float dt = (float)std::atof(argv[4]);
    const int frames = std::atoi(argv[5]);
    const bool priv = argc > 7 && std::string(argv[7]) == "private";
    if (recs.empty() || n_ctx < 1) return 4;

    const Result one = run(recs, 1, false, dt, frames, false);
    const Result split = run(recs, n_ctx, spread, dt, frames, priv);
    bool same = true;
    for (size_t e = 0; e < recs.size(); ++e)
        same = same && one.counts[e] == split.counts[e] && one.states[e] == split.states[e] &&
               one.verts[e].size() == split.verts[e].size() &&
               std::memcmp(one.verts[e].data(), split.verts[e].data(), one.verts[e].size() * sizeof(float)) == 0;
    FILE* out = std::fopen(argv[6], "wb");
    if (!out) return 5;
    for (size_t e = 0; e < recs.size(); ++e)
    {
        std::fwrite(&split.counts[e], sizeof(uint32_t), 1, out);
        std::fwrite(split.states[e].data(), sizeof(uint32_t), split.states[e].size(), out);
    }
    std::fclose(out);
    uint64_t total = 0;
    for (uint32_t n : one.counts) total += n;
    std::printf("%zu emitters, %d context(s) %s, dt %g ms, %d frames, %llu live particles: split run %s the single-context run\n", recs.size(),
                n_ctx, spread ? "on separate devices" : "on one device", dt, frames, (unsigned long long)total, same ? "EQUALS" : "DIFFERS FROM");
    return same ? 0 : 1;
}
