// facade_driver.cpp -- drives the C++ ParticleEmitter facade the way ParticlesSample drives the reference
// class (create from a .particle file, attach to a node, start, update/draw per frame) and dumps the result
// for the Python test to compare with the oracle.
//   facade_driver <file.particle> <frames> <seed> <out.bin> [emit_once_n] [props]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../tractor3d_b200/host/ParticleEmitter.h"

using namespace tractor;

int main(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s file frames seed out [burst]\n", argv[0]); return 2; }
    const int frames = std::atoi(argv[2]);
    const uint64_t seed = std::strtoull(argv[3], nullptr, 10);
    const unsigned burst = argc > 5 ? (unsigned)std::atoi(argv[5]) : 0;

    ParticleEmitter* e = nullptr;
    if (argc > 6 && std::string(argv[6]) == "props")
    {
        // the way src/scene/SceneLoader.cpp:318-324 gets its emitter: a parsed Properties tree, one namespace of it handed over
        Properties* file = Properties::create(argv[1]);
        if (!file) return 3;
        e = ParticleEmitter::create(file->getNamespace().length() > 0 ? file : file->getNextNamespace());
        delete file;
        if (!e->getTexture() || e->getTexture()->getWidth() == 0) return 9; // PE.cpp:255-259
    }
    else
        e = ParticleEmitter::create(argv[1]);
    if (!e) return 3;
    // A missing file must come back as nullptr, not as an exit (built with GP_ERRORS_AS_WARNINGS).
    if (ParticleEmitter::create("/nonexistent/none.particle") != nullptr) return 4;

    Node node, cameraNode;
    Scene scene;
    Camera camera;
    Matrix cam;
    cam.m[12] = 0.22f; cam.m[13] = 2.7f; cam.m[14] = 13.3f;
    cameraNode.setWorldMatrix(cam);
    camera.setNode(&cameraNode);
    scene.setActiveCamera(&camera);
    node.setScene(&scene);
    node.setDrawable(e);

    e->setRandomSource(T3D_RNG_DEVICE, seed);
    if (e->isActive()) return 5; // not started, no particles (PE.cpp:277-284)
    if (burst) e->emitOnce(burst);
    e->start();
    unsigned draws = 0;
    for (int f = 0; f < frames; ++f)
    {
        Matrix w;
        w.m[12] = 0.05f * (float)f; // the node drifts; spawn positions follow (PE.cpp:354)
        node.setWorldMatrix(w);
        e->update(1000.0f / 60.0f);
        draws += e->draw();
    }
    const unsigned count = e->getParticlesCount();
    std::vector<uint32_t> state((size_t)T3D_NUM_COLUMNS * (count ? count : 1));
    uint32_t n = 0;
    if (t3d_emitter_download_state(e->handle(), state.data(), count ? count : 1, &n) != T3D_OK || n != count) return 6;
    unsigned quads = 0;
    e->getVertexStream(&quads);
    std::vector<float> verts((size_t)T3D_FLOATS_PER_QUAD * (quads ? quads : 1));
    uint32_t q = 0;
    if (t3d_emitter_download_vertices(e->handle(), verts.data(), quads, &q) != T3D_OK) return 7;

    FILE* out = std::fopen(argv[4], "wb");
    if (!out) return 8;
    const uint32_t hdr[4] = { count, quads, draws, e->getSpriteFrameCount() };
    std::fwrite(hdr, sizeof(hdr), 1, out);
    std::fwrite(state.data(), sizeof(uint32_t), (size_t)T3D_NUM_COLUMNS * count, out);
    std::fwrite(verts.data(), sizeof(float), (size_t)T3D_FLOATS_PER_QUAD * quads, out);
    std::fclose(out);
    std::printf("count %u quads %u draws %u rate %u energy %ld..%ld sprite %ux%u\n", count, quads, draws, e->getEmissionRate(),
                e->getEnergyMin(), e->getEnergyMax(), e->getSpriteWidth(), e->getSpriteHeight());
    node.setDrawable(nullptr);
    e->release();
    return 0;
}
