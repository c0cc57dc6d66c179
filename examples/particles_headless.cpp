// particles_headless.cpp -- the browser sample's particle loop (reference samples/browser/src/ParticlesSample.cpp:
// initialize :1077-1204, update :863-898, render :901-921) without a window: load a .particle description, attach the
// emitter to a node under a scene with an active camera, start it, and run update()/draw() at 60 Hz.
// The only lines that differ from the sample are the headless scaffolding and the last paragraph, where the vertex
// stream is taken from the device instead of being handed to GL.
//
//   g++ -std=c++17 -O2 -DGP_ERRORS_AS_WARNINGS examples/particles_headless.cpp tractor3d_b200/host/ParticleEmitter.cpp \
//       -Ltractor3d_b200/csrc -lt3d_particles -Wl,-rpath,$PWD/tractor3d_b200/csrc -o particles_headless
//   ./particles_headless path/to/fire.particle [frames]
#include <cstdio>
#include <cstdlib>

#include "../tractor3d_b200/host/ParticleEmitter.h"

using namespace tractor;

int main(int argc, char** argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s file.particle [frames]\n", argv[0]);
        return 2;
    }
    const int frames = argc > 2 ? std::atoi(argv[2]) : 600;

    ParticleEmitter* emitter = ParticleEmitter::create(argv[1]);   // ParticlesSample.cpp:1187
    if (!emitter) return 1;

    Scene scene;
    Node cameraNode, particleNode;
    Camera camera;
    Matrix cam;                       // editor camera of the sample: translation only (ParticlesSample.cpp:1102-1115)
    cam.m[12] = 0.22f; cam.m[13] = 2.7f; cam.m[14] = 13.3f;
    cameraNode.setWorldMatrix(cam);
    camera.setNode(&cameraNode);
    scene.setActiveCamera(&camera);
    particleNode.setScene(&scene);
    particleNode.setDrawable(emitter);                             // ParticlesSample.cpp:1192

    emitter->start();                                              // ParticlesSample.cpp:1203

    for (int f = 0; f < frames; ++f)
    {
        emitter->update(1000.0f / 60.0f);                          // ParticlesSample.cpp:897-898
        emitter->draw();                                           // ParticlesSample.cpp:916-921
        if (f % 100 == 99)
        {
            const unsigned n = emitter->getParticlesCount();       // ParticlesSample.cpp:1206-1220 shows this number
            std::printf("frame %4d: %u live particles\n", f + 1, n);
        }
    }
    ParticleEmitter::sync();

    unsigned quads = 0;
    const t3d_sprite_vertex* stream = emitter->getVertexStream(&quads);
    std::printf("%u quads in the device-resident vertex stream at %p (4 SpriteVertex each, SpriteBatch.h:360-380)\n", quads,
                (const void*)stream);
    particleNode.setDrawable(nullptr);
    emitter->release();
    return 0;
}
