// engine_caller.cpp -- application code written against the ENGINE's headers only: what src/scene/SceneLoader.cpp:318-324
// does with a `particle` namespace, and every ParticleEmitter member samples/browser/src/ParticlesSample.cpp calls
// (:323-336 blend modes, :437-777 the editor's getters and setters, :834-852 getTexture + setSpriteFrameCoords,
// :897-921 update / draw, :1081-1095 create / release, :1138-1204 the getters that fill the form).
//
// It is compiled twice from this one file (tests/test_engine_headers.py):
//   * against the reference's own graphics/ParticleEmitter.h, linked with the unmodified reference (oracle/_ref);
//   * against tractor3d_b200/host/engine/graphics/ParticleEmitter.h, i.e. the B200 backend's class, with
//     -DT3D_WITH_ENGINE_HEADERS, the engine's math / Ref / Properties objects and libt3d_particles.so.
// Everything it prints is independent of the random stream, so the two programs must print the same text.
#include "pch.h"

#include <cstdio>
#include <string>

#include "graphics/ParticleEmitter.h"
#include "scene/Node.h"
#include "scene/Properties.h"
#include "scene/Scene.h"

using namespace tractor;

static const char* blendName(ParticleEmitter::BlendMode m) // ParticlesSample.cpp:323-336
{
    switch (m)
    {
        case ParticleEmitter::BLEND_NONE: return "BLEND_NONE";
        case ParticleEmitter::BLEND_ALPHA: return "BLEND_ALPHA";
        case ParticleEmitter::BLEND_ADDITIVE: return "BLEND_ADDITIVE";
        case ParticleEmitter::BLEND_MULTIPLIED: return "BLEND_MULTIPLIED";
    }
    return "?";
}

static void form(const char* tag, ParticleEmitter* e) // ParticlesSample.cpp:1138-1204
{
    std::printf("[%s] color %g %g %g %g / %g %g %g %g -> %g %g %g %g / %g %g %g %g\n", tag, e->getColorStart().x, e->getColorStart().y,
                e->getColorStart().z, e->getColorStart().w, e->getColorStartVariance().x, e->getColorStartVariance().y,
                e->getColorStartVariance().z, e->getColorStartVariance().w, e->getColorEnd().x, e->getColorEnd().y, e->getColorEnd().z,
                e->getColorEnd().w, e->getColorEndVariance().x, e->getColorEndVariance().y, e->getColorEndVariance().z,
                e->getColorEndVariance().w);
    std::printf("[%s] size %g %g %g %g energy %ld %ld rate %u max %u ellipsoid %d\n", tag, e->getSizeStartMin(), e->getSizeStartMax(),
                e->getSizeEndMin(), e->getSizeEndMax(), e->getEnergyMin(), e->getEnergyMax(), e->getEmissionRate(), e->getParticleCountMax(),
                (int)e->isEllipsoid());
    const Vector3 p = e->getPosition(), pv = e->getPositionVariance(), v = e->getVelocity(), vv = e->getVelocityVariance();
    const Vector3 a = e->getAcceleration(), av = e->getAccelerationVariance(), ax = e->getRotationAxis(), axv = e->getRotationAxisVariance();
    std::printf("[%s] pos %g %g %g / %g %g %g vel %g %g %g / %g %g %g acc %g %g %g / %g %g %g\n", tag, p.x, p.y, p.z, pv.x, pv.y, pv.z, v.x,
                v.y, v.z, vv.x, vv.y, vv.z, a.x, a.y, a.z, av.x, av.y, av.z);
    std::printf("[%s] spin %g %g rot %g %g axis %g %g %g / %g %g %g orbit %d %d %d\n", tag, e->getRotationPerParticleSpeedMin(),
                e->getRotationPerParticleSpeedMax(), e->getRotationSpeedMin(), e->getRotationSpeedMax(), ax.x, ax.y, ax.z, axv.x, axv.y, axv.z,
                (int)e->getOrbitPosition(), (int)e->getOrbitVelocity(), (int)e->getOrbitAcceleration());
    std::printf("[%s] sprite %ux%u frames %u animated %d looped %d offset %d duration %ld blend %s\n", tag, e->getSpriteWidth(),
                e->getSpriteHeight(), e->getSpriteFrameCount(), (int)e->isSpriteAnimated(), (int)e->isSpriteLooped(),
                e->getSpriteFrameRandomOffset(), e->getSpriteFrameDuration(), blendName(e->getBlendMode()));
    Texture* t = e->getTexture();
    std::printf("[%s] texture %ux%u started %d active %d count %u\n", tag, t ? t->getWidth() : 0, t ? t->getHeight() : 0, (int)e->isStarted(),
                (int)e->isActive(), e->getParticlesCount());
}

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        std::fprintf(stderr, "usage: %s scene-like.properties other-texture.png\n", argv[0]);
        return 2;
    }
    // ---- SceneLoader.cpp:318-324: a parsed tree holds a `particle` namespace; attach an emitter made from it to a node ----
    Properties* file = Properties::create(argv[1]);
    if (!file) return 3;
    Properties* p = file->getNamespace().length() > 0 ? file : file->getNextNamespace();
    Node node, cameraNode;
    Scene scene;
    Camera camera;
    camera._node = &cameraNode;
    scene._camera = &camera;
    node._scene = &scene;
    ParticleEmitter* particleEmitter = ParticleEmitter::create(p);
    if (!particleEmitter) return 4;
    node.setDrawable(particleEmitter);
    SAFE_RELEASE(particleEmitter);
    ParticleEmitter* e = dynamic_cast<ParticleEmitter*>(node._drawable);
    form("loaded", e);

    // ---- ParticlesSample.cpp:437-777: the editor's setters, each fed from the matching getters ----
    Vector4 c = e->getColorStart();
    c.x = 0.25f;
    e->setColor(c, e->getColorStartVariance(), e->getColorEnd(), e->getColorEndVariance());
    e->setSize(0.5f, e->getSizeStartMax(), e->getSizeEndMin(), 3.0f);
    e->setEnergy(400L, 400L); // a fixed lifetime: nothing below depends on the random stream
    e->setEmissionRate(250);
    Vector3 v = e->getPosition();
    v.y = 1.5f;
    e->setPosition(v, e->getPositionVariance());
    e->setVelocity(e->getVelocity(), Vector3(0.5f, 0.25f, 0.125f));
    e->setAcceleration(Vector3(0.0f, -9.0f, 0.0f), e->getAccelerationVariance());
    e->setRotationPerParticle(-2.0f, e->getRotationPerParticleSpeedMax());
    e->setRotation(e->getRotationSpeedMin(), 0.0f, Vector3(0.0f, 1.0f, 0.0f), e->getRotationAxisVariance());
    e->setOrbit(true, false, true);
    e->setEllipsoid(!e->isEllipsoid());
    e->setBlendMode(ParticleEmitter::BLEND_MULTIPLIED);
    e->setSpriteAnimated(true);
    e->setSpriteLooped(false);
    e->setSpriteFrameRandomOffset(2);
    e->setSpriteFrameDuration(75L);
    form("edited", e);

    // ---- :834-852: the sprite sheet is re-cut from the texture's own size ----
    Texture* texture = e->getTexture();
    e->setSpriteFrameCoords(6, (int)texture->getWidth() / 3, (int)texture->getHeight() / 2);
    form("recut", e);

    // ---- :772-777, :897-921: bursts, frames, and the count the loop leaves behind ----
    e->emitOnce(40);
    std::printf("burst: count %u active %d\n", e->getParticlesCount(), (int)e->isActive());
    e->setParticleCountMax(60); // PE.h:237: the clamp of PE.cpp:293-296 now uses the lower value
    e->emitOnce(100);
    std::printf("clamped: count %u max %u\n", e->getParticlesCount(), e->getParticleCountMax());
    unsigned draws = 0;
    for (int f = 0; f < 10; ++f)
    {
        node._world.m[12] = 0.1f * (float)f;
        e->update(1000.0f / 60.0f);
        draws += e->draw();
    }
    std::printf("10 frames: count %u draws %u\n", e->getParticlesCount(), draws);
    for (int f = 0; f < 20; ++f) e->update(1000.0f / 60.0f); // 400 ms lifetime: all gone after 24 frames
    std::printf("30 frames: count %u active %d draw %u\n", e->getParticlesCount(), (int)e->isActive(), e->draw());
    e->start();
    e->update(20.0f); // 250 particles per second: 5 emitted
    std::printf("started: count %u started %d active %d\n", e->getParticlesCount(), (int)e->isStarted(), (int)e->isActive());
    e->stop();

    // ---- PE.cpp:214-252: a new texture resets the sprite table to one frame covering it ----
    e->setTexture(argv[2], ParticleEmitter::BLEND_ADDITIVE);
    form("retextured", e);

    node.setDrawable(nullptr);
    delete file;
    return 0;
}
