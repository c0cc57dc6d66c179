// ParticleEmitter.h -- drop-in for tractor::ParticleEmitter (reference
// Tractor3Dlib/include/graphics/ParticleEmitter.h:152-734) whose particles live in HBM and whose
// emitOnce / update / draw run as sm_100a kernels behind the C ABI of include/t3d_particles.h.
//
// Same names, same argument meaning, same units (milliseconds, particles per second), same ownership
// convention (Ref: created with one reference, release() to drop it), same error behaviour (GP_ERROR /
// GP_WARN, create*() return nullptr on failure).  What differs, by design:
//   * there is no SpriteBatch: the texture is only ever asked for its size (PE.cpp:244-247; getTexture() hands back
//     the object setTexture() was given) and the vertex stream stays on the device -- getVertexStream() hands the
//     renderer the device pointer where the reference handed MeshBatch a client-side array;
//   * update()/draw() enqueue; calls on many emitters coalesce into one launch per phase.  flush()/sync()
//     on the shared context (or any getter that needs a result) make them run;
//   * setParticleCountMax() cannot grow the emitter beyond the size it was created with (the reference only rewrites
//     the field and would overrun its array, PE.h:237): growth is refused with GP_WARN; lowering it works as there.
// Builds against compat.h's stand-ins or, with -DT3D_WITH_ENGINE_HEADERS, against the engine's own headers.
#pragma once

#include <cstdint>
#include <string>

#include "../../include/t3d_particles.h"
#include "compat.h"

namespace tractor
{

class ParticleEmitter : public Ref, public Drawable
{
    friend class Node;

  public:
    enum BlendMode { BLEND_NONE, BLEND_ALPHA, BLEND_ADDITIVE, BLEND_MULTIPLIED };

    // One context per GPU, shared by every emitter of the process (created on first use on device 0 unless
    // setDevice() was called first).
    static void setDevice(int device);
    static t3d_ctx* context();
    static void flush();
    static void sync();
    static void setRotationMode(int t3d_rot_mode);

    static ParticleEmitter* create(const std::string& url);                                    // PE.cpp:76
    static ParticleEmitter* create(Properties* properties);                                    // PE.cpp:92 (SceneLoader.cpp:320)
    static ParticleEmitter* create(const std::string& texturePath, BlendMode blendMode,
                                   unsigned int particleCountMax);                             // PE.cpp:43
    // No engine Texture here: the size is all the path needs.
    static ParticleEmitter* create(unsigned int textureWidth, unsigned int textureHeight, BlendMode blendMode,
                                   unsigned int particleCountMax);

    void setTexture(const std::string& texturePath, BlendMode blendMode);                      // PE.cpp:214
    void setTexture(Texture* texture, BlendMode blendMode);                                    // PE.cpp:229
    Texture* getTexture() const;                                                               // PE.cpp:255
    void setParticleCountMax(unsigned int max);
    unsigned int getParticleCountMax() const;
    void setEmissionRate(unsigned int rate);
    unsigned int getEmissionRate() const;
    void start();
    void stop();
    bool isStarted() const;
    bool isActive() const;
    void emitOnce(unsigned int particleCount);
    unsigned int getParticlesCount() const;
    void setEllipsoid(bool ellipsoid);
    bool isEllipsoid() const;
    void setSize(float startMin, float startMax, float endMin, float endMax);
    float getSizeStartMin() const;
    float getSizeStartMax() const;
    float getSizeEndMin() const;
    float getSizeEndMax() const;
    void setColor(const Vector4& start, const Vector4& startVariance, const Vector4& end, const Vector4& endVariance);
    Vector4 getColorStart() const;
    Vector4 getColorStartVariance() const;
    Vector4 getColorEnd() const;
    Vector4 getColorEndVariance() const;
    void setEnergy(long energyMin, long energyMax);
    long getEnergyMin() const;
    long getEnergyMax() const;
    void setPosition(const Vector3& position, const Vector3& positionVariance);
    Vector3 getPosition() const;
    Vector3 getPositionVariance() const;
    void setVelocity(const Vector3& velocity, const Vector3& velocityVariance);
    Vector3 getVelocity() const;
    Vector3 getVelocityVariance() const;
    void setAcceleration(const Vector3& acceleration, const Vector3& accelerationVariance);
    Vector3 getAcceleration() const;
    Vector3 getAccelerationVariance() const;
    void setRotationPerParticle(float speedMin, float speedMax);
    float getRotationPerParticleSpeedMin() const;
    float getRotationPerParticleSpeedMax() const;
    void setRotation(float speedMin, float speedMax, const Vector3& axis, const Vector3& axisVariance);
    float getRotationSpeedMin() const;
    float getRotationSpeedMax() const;
    Vector3 getRotationAxis() const;
    Vector3 getRotationAxisVariance() const;
    void setSpriteAnimated(bool animated);
    bool isSpriteAnimated() const;
    void setSpriteLooped(bool looped);
    bool isSpriteLooped() const;
    void setSpriteFrameRandomOffset(int maxOffset);
    int getSpriteFrameRandomOffset() const;
    void setSpriteFrameDuration(long duration);
    long getSpriteFrameDuration() const;
    unsigned int getSpriteWidth() const;
    unsigned int getSpriteHeight() const;
    void setSpriteTexCoords(unsigned int frameCount, float* texCoords);
    void setSpriteFrameCoords(unsigned int frameCount, Rectangle* frameCoords);
    void setSpriteFrameCoords(unsigned int frameCount, int width, int height);
    unsigned int getSpriteFrameCount() const;
    void setOrbit(bool orbitPosition, bool orbitVelocity, bool orbitAcceleration);
    bool getOrbitPosition() const;
    bool getOrbitVelocity() const;
    bool getOrbitAcceleration() const;
    void setBlendMode(BlendMode blendMode);
    BlendMode getBlendMode() const;
    void update(float elapsedTime);
    unsigned int draw(bool wireframe = false) override;

    // ---- additions (not in the reference) ----
    // Where the draws come from: T3D_RNG_LIBC (default: rand(), like the reference), _STREAM, _DEVICE.
    void setRandomSource(int t3d_rng_mode, uint64_t seed = 0);
    void feedDraws(const int32_t* draws, size_t n);
    // Device-resident vertex stream of the last draw(): 4 SpriteVertex per quad, live-array order.
    const t3d_sprite_vertex* getVertexStream(unsigned int* quadCount) const;
    t3d_emitter* handle() const { return _h; }

  protected:
    ParticleEmitter() = default;
    ~ParticleEmitter() override;
    Drawable* clone(NodeCloneContext& context) override;                                       // PE.cpp:891
    ParticleEmitter(const ParticleEmitter&) = delete;
    ParticleEmitter& operator=(const ParticleEmitter&) = delete;

  private:
    t3d_emitter_params params() const;
    void apply(const t3d_emitter_params& p);
    bool pushTransforms();
    static BlendMode getBlendModeFromString(const std::string& str);
    static ParticleEmitter* create(Texture* texture, BlendMode blendMode, unsigned int particleCountMax); // PE.cpp:63 (PE.h:760)

    t3d_emitter* _h{ nullptr };
    Texture* _texture{ nullptr }; // what the reference keeps inside its SpriteBatch's sampler (PE.cpp:255-259)
};

} // namespace tractor
