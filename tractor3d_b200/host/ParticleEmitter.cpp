// ParticleEmitter.cpp -- the reference's ParticleEmitter interface over the C ABI (see ParticleEmitter.h).
// Each method cites the reference member it stands for; PE.cpp = Tractor3Dlib/src/graphics/ParticleEmitter.cpp.
#include "ParticleEmitter.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

namespace tractor
{

namespace
{
t3d_ctx* g_ctx = nullptr;
int g_device = 0;

void copy3(float* d, const Vector3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void copy4(float* d, const Vector4& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
Vector3 vec3(const float* s) { return Vector3(s[0], s[1], s[2]); }
Vector4 vec4(const float* s) { return Vector4(s[0], s[1], s[2], s[3]); }
} // namespace

void ParticleEmitter::setDevice(int device) { g_device = device; }

t3d_ctx* ParticleEmitter::context()
{
    if (!g_ctx && t3d_ctx_create(g_device, nullptr, &g_ctx) != T3D_OK)
    {
        GP_ERROR("Failed to create the particle backend context: %s", t3d_last_error());
        return nullptr;
    }
    return g_ctx;
}

void ParticleEmitter::flush() { if (g_ctx) t3d_ctx_flush(g_ctx); }
void ParticleEmitter::sync() { if (g_ctx) t3d_ctx_sync(g_ctx); }
void ParticleEmitter::setRotationMode(int mode) { if (context()) t3d_ctx_set_rotation_mode(g_ctx, mode); }

ParticleEmitter::~ParticleEmitter()
{
    if (_h) t3d_emitter_destroy(_h); // PE.cpp:35-40: the emitter owns its particle storage
    if (_texture) _texture->release();
}

// PE.cpp:76-89
ParticleEmitter* ParticleEmitter::create(const std::string& url)
{
    auto properties = std::unique_ptr<Properties>(Properties::create(url));
    if (!properties)
    {
        GP_ERROR("Failed to create particle emitter from file.");
        return nullptr;
    }
    return create(properties->getNamespace().length() > 0 ? properties.get() : properties->getNextNamespace());
}

// PE.cpp:92-211, getter by getter: whatever Properties class this is built against (the engine's, or compat.h's view of
// the library's own parser) does the conversions.
ParticleEmitter* ParticleEmitter::create(Properties* properties)
{
    if (!properties || properties->getNamespace() != "particle")
    {
        GP_ERROR("Properties object must be non-null and have namespace equal to 'particle'.");
        return nullptr;
    }
    properties->rewind();
    Properties* sprite = properties->getNextNamespace();
    if (!sprite || sprite->getNamespace() != "sprite")
    {
        GP_ERROR("Failed to load particle emitter: required namespace 'sprite' is missing.");
        return nullptr;
    }
    std::string texturePath;
    if (!sprite->getPath("path", texturePath))
    {
        GP_ERROR("Failed to load particle emitter: required image file path ('path') is missing.");
        return nullptr;
    }
    std::string blendModeString = sprite->getString("blendMode");
    if (blendModeString.empty()) blendModeString = sprite->getString("blending"); // the old naming
    const BlendMode blendMode = getBlendModeFromString(blendModeString);
    const int spriteWidth = sprite->getInt("width");
    const int spriteHeight = sprite->getInt("height");
    const bool spriteAnimated = sprite->getBool("animated");
    const bool spriteLooped = sprite->getBool("looped");
    const int spriteFrameCount = sprite->getInt("frameCount");
    const int spriteFrameRandomOffset = std::min(sprite->getInt("frameRandomOffset"), spriteFrameCount);
    const float spriteFrameDuration = sprite->getFloat("frameDuration");

    unsigned int particleCountMax = (unsigned int)properties->getInt("particleCountMax");
    if (particleCountMax == 0) particleCountMax = 100; // PARTICLE_COUNT_MAX, PE.cpp:12
    unsigned int emissionRate = (unsigned int)properties->getInt("emissionRate");
    if (emissionRate == 0) emissionRate = 10;          // PARTICLE_EMISSION_RATE, PE.cpp:13
    const bool ellipsoid = properties->getBool("ellipsoid");
    const float sizeStartMin = properties->getFloat("sizeStartMin");
    const float sizeStartMax = properties->getFloat("sizeStartMax");
    const float sizeEndMin = properties->getFloat("sizeEndMin");
    const float sizeEndMax = properties->getFloat("sizeEndMax");
    const long energyMin = properties->getLong("energyMin");
    const long energyMax = properties->getLong("energyMax");
    Vector4 colorStart, colorStartVar, colorEnd, colorEndVar;
    properties->getVector4("colorStart", &colorStart);
    properties->getVector4("colorStartVar", &colorStartVar);
    properties->getVector4("colorEnd", &colorEnd);
    properties->getVector4("colorEndVar", &colorEndVar);
    Vector3 position, positionVar, velocity, velocityVar, acceleration, accelerationVar, rotationAxis, rotationAxisVar;
    properties->getVector3("position", &position);
    properties->getVector3("positionVar", &positionVar);
    properties->getVector3("velocity", &velocity);
    properties->getVector3("velocityVar", &velocityVar);
    properties->getVector3("acceleration", &acceleration);
    properties->getVector3("accelerationVar", &accelerationVar);
    const float rotationPerParticleSpeedMin = properties->getFloat("rotationPerParticleSpeedMin");
    const float rotationPerParticleSpeedMax = properties->getFloat("rotationPerParticleSpeedMax");
    const float rotationSpeedMin = properties->getFloat("rotationSpeedMin");
    const float rotationSpeedMax = properties->getFloat("rotationSpeedMax");
    properties->getVector3("rotationAxis", &rotationAxis);
    properties->getVector3("rotationAxisVar", &rotationAxisVar);
    const bool orbitPosition = properties->getBool("orbitPosition");
    const bool orbitVelocity = properties->getBool("orbitVelocity");
    const bool orbitAcceleration = properties->getBool("orbitAcceleration");

    ParticleEmitter* emitter = ParticleEmitter::create(texturePath, blendMode, particleCountMax);
    if (!emitter)
    {
        GP_ERROR("Failed to create particle emitter.");
        return nullptr;
    }
    emitter->setEmissionRate(emissionRate);
    emitter->setEllipsoid(ellipsoid);
    emitter->setSize(sizeStartMin, sizeStartMax, sizeEndMin, sizeEndMax);
    emitter->setEnergy(energyMin, energyMax);
    emitter->setColor(colorStart, colorStartVar, colorEnd, colorEndVar);
    emitter->setPosition(position, positionVar);
    emitter->setVelocity(velocity, velocityVar);
    emitter->setAcceleration(acceleration, accelerationVar);
    emitter->setRotationPerParticle(rotationPerParticleSpeedMin, rotationPerParticleSpeedMax);
    emitter->setRotation(rotationSpeedMin, rotationSpeedMax, rotationAxis, rotationAxisVar);
    emitter->setSpriteAnimated(spriteAnimated);
    emitter->setSpriteLooped(spriteLooped);
    emitter->setSpriteFrameRandomOffset(spriteFrameRandomOffset);
    emitter->setSpriteFrameDuration((long)spriteFrameDuration);
    emitter->setSpriteFrameCoords(spriteFrameCount, spriteWidth, spriteHeight);
    emitter->setOrbit(orbitPosition, orbitVelocity, orbitAcceleration);
    return emitter;
}

// PE.cpp:43-60: the texture is loaded (here: its header is read) and handed on.
ParticleEmitter* ParticleEmitter::create(const std::string& texturePath, BlendMode blendMode, unsigned int particleCountMax)
{
    Texture* texture = Texture::create(texturePath, true);
    if (!texture)
    {
        GP_ERROR("Failed to create texture for particle emitter.");
        return nullptr;
    }
    ParticleEmitter* emitter = ParticleEmitter::create(texture, blendMode, particleCountMax);
    texture->release();
    return emitter;
}

// PE.cpp:63-73
ParticleEmitter* ParticleEmitter::create(Texture* texture, BlendMode blendMode, unsigned int particleCountMax)
{
    ParticleEmitter* emitter = create(texture->getWidth(), texture->getHeight(), blendMode, particleCountMax);
    if (emitter)
    {
        texture->addRef();
        emitter->_texture = texture;
    }
    return emitter;
}

// The same without a Texture object: the size is all the path needs (getTexture() then returns nullptr).
ParticleEmitter* ParticleEmitter::create(unsigned int textureWidth, unsigned int textureHeight, BlendMode blendMode,
                                         unsigned int particleCountMax)
{
    t3d_ctx* ctx = context();
    if (!ctx) return nullptr;
    t3d_emitter_params p;
    t3d_emitter_params_default(&p);
    p.particle_count_max = particleCountMax;
    p.texture_width = (float)textureWidth;
    p.texture_height = (float)textureHeight;
    p.blend_mode = (int32_t)blendMode;
    t3d_emitter* h = nullptr;
    if (t3d_emitter_create(ctx, &p, &h) != T3D_OK)
    {
        GP_ERROR("Failed to create particle emitter: %s", t3d_last_error());
        return nullptr;
    }
    ParticleEmitter* e = new ParticleEmitter();
    e->_h = h;
    return e;
}

// PE.cpp:214-226
void ParticleEmitter::setTexture(const std::string& texturePath, BlendMode blendMode)
{
    Texture* texture = Texture::create(texturePath, true);
    if (texture)
    {
        setTexture(texture, blendMode);
        texture->release();
    }
    else
    {
        GP_WARN("Failed set new texture on particle emitter: %s", texturePath.c_str());
    }
}

// PE.cpp:229-252: the new texture's size re-derives the texel ratios and the sprite table goes back to one frame
// covering the whole texture.  (New reference taken before the old one is dropped: it may be the same texture.)
void ParticleEmitter::setTexture(Texture* texture, BlendMode blendMode)
{
    if (t3d_emitter_set_texture_size(_h, texture->getWidth(), texture->getHeight(), (int)blendMode) != T3D_OK)
    {
        GP_ERROR("setTexture: %s", t3d_last_error());
        return;
    }
    texture->addRef();
    if (_texture) _texture->release();
    _texture = texture;
}

Texture* ParticleEmitter::getTexture() const { return _texture; } // PE.cpp:255-259

t3d_emitter_params ParticleEmitter::params() const
{
    t3d_emitter_params p;
    t3d_emitter_get_params(_h, &p);
    return p;
}

void ParticleEmitter::apply(const t3d_emitter_params& p)
{
    if (t3d_emitter_set_params(_h, &p) != T3D_OK) GP_WARN("%s", t3d_last_error());
}

void ParticleEmitter::setParticleCountMax(unsigned int max)
{
    t3d_emitter_params p = params();
    // PE.h:237 rewrites the field; the storage keeps the size it was created with, so the field may go down (and back up
    // to that size): the library refuses anything beyond it
    p.particle_count_max = max;
    if (t3d_emitter_set_params(_h, &p) != T3D_OK) GP_WARN("setParticleCountMax(%u): %s", max, t3d_last_error());
}
unsigned int ParticleEmitter::getParticleCountMax() const { return params().particle_count_max; }

void ParticleEmitter::setEmissionRate(unsigned int rate) // PE.cpp:262-267
{
    if (t3d_emitter_set_emission_rate(_h, rate) != T3D_OK) GP_ERROR("%s", t3d_last_error());
}
unsigned int ParticleEmitter::getEmissionRate() const { return params().emission_rate; }

void ParticleEmitter::start() { t3d_emitter_start(_h); } // PE.cpp:270-274
void ParticleEmitter::stop() { t3d_emitter_stop(_h); }   // PE.h:272
bool ParticleEmitter::isStarted() const { int v = 0; t3d_emitter_is_started(_h, &v); return v != 0; }

bool ParticleEmitter::isActive() const // PE.cpp:277-284
{
    const_cast<ParticleEmitter*>(this)->pushTransforms();
    int v = 0;
    t3d_emitter_is_active(_h, &v);
    return v != 0;
}

// The reference reads its node (and, for draw, the scene's active camera node) at call time: hand them over.
bool ParticleEmitter::pushTransforms()
{
    if (!_node) return false;
    t3d_emitter_set_node_world(_h, _node->getWorldMatrix().m);
    Scene* scene = _node->getScene();
    Camera* cam = scene ? scene->getActiveCamera() : nullptr;
    if (cam && cam->getNode()) t3d_emitter_set_camera_world(_h, cam->getNode()->getWorldMatrix().m);
    return true;
}

void ParticleEmitter::emitOnce(unsigned int particleCount) // PE.cpp:287-369
{
    pushTransforms();
    if (t3d_emitter_emit_once(_h, particleCount) != T3D_OK) GP_ERROR("emitOnce: %s", t3d_last_error());
}

unsigned int ParticleEmitter::getParticlesCount() const
{
    uint32_t n = 0;
    t3d_emitter_get_count(_h, &n);
    return n;
}

void ParticleEmitter::setEllipsoid(bool v) { t3d_emitter_params p = params(); p.ellipsoid = v; apply(p); }
bool ParticleEmitter::isEllipsoid() const { return params().ellipsoid != 0; }

void ParticleEmitter::setSize(float startMin, float startMax, float endMin, float endMax) // PE.cpp:372-378
{
    t3d_emitter_params p = params();
    p.size_start_min = startMin; p.size_start_max = startMax; p.size_end_min = endMin; p.size_end_max = endMax;
    apply(p);
}
float ParticleEmitter::getSizeStartMin() const { return params().size_start_min; }
float ParticleEmitter::getSizeStartMax() const { return params().size_start_max; }
float ParticleEmitter::getSizeEndMin() const { return params().size_end_min; }
float ParticleEmitter::getSizeEndMax() const { return params().size_end_max; }

void ParticleEmitter::setColor(const Vector4& s, const Vector4& sv, const Vector4& e, const Vector4& ev) // PE.cpp:388-397
{
    t3d_emitter_params p = params();
    copy4(p.color_start, s); copy4(p.color_start_var, sv); copy4(p.color_end, e); copy4(p.color_end_var, ev);
    apply(p);
}
Vector4 ParticleEmitter::getColorStart() const { return vec4(params().color_start); }
Vector4 ParticleEmitter::getColorStartVariance() const { return vec4(params().color_start_var); }
Vector4 ParticleEmitter::getColorEnd() const { return vec4(params().color_end); }
Vector4 ParticleEmitter::getColorEndVariance() const { return vec4(params().color_end_var); }

void ParticleEmitter::setEnergy(long energyMin, long energyMax) // PE.cpp:381-385: long arguments, float members
{
    t3d_emitter_params p = params();
    p.energy_min = (float)energyMin; p.energy_max = (float)energyMax;
    apply(p);
}
long ParticleEmitter::getEnergyMin() const { return (long)params().energy_min; } // PE.h:427
long ParticleEmitter::getEnergyMax() const { return (long)params().energy_max; }

void ParticleEmitter::setPosition(const Vector3& v, const Vector3& var)
{
    t3d_emitter_params p = params(); copy3(p.position, v); copy3(p.position_var, var); apply(p);
}
Vector3 ParticleEmitter::getPosition() const { return vec3(params().position); }
Vector3 ParticleEmitter::getPositionVariance() const { return vec3(params().position_var); }
void ParticleEmitter::setVelocity(const Vector3& v, const Vector3& var)
{
    t3d_emitter_params p = params(); copy3(p.velocity, v); copy3(p.velocity_var, var); apply(p);
}
Vector3 ParticleEmitter::getVelocity() const { return vec3(params().velocity); }
Vector3 ParticleEmitter::getVelocityVariance() const { return vec3(params().velocity_var); }
void ParticleEmitter::setAcceleration(const Vector3& v, const Vector3& var)
{
    t3d_emitter_params p = params(); copy3(p.acceleration, v); copy3(p.acceleration_var, var); apply(p);
}
Vector3 ParticleEmitter::getAcceleration() const { return vec3(params().acceleration); }
Vector3 ParticleEmitter::getAccelerationVariance() const { return vec3(params().acceleration_var); }

void ParticleEmitter::setRotationPerParticle(float speedMin, float speedMax)
{
    t3d_emitter_params p = params();
    p.rotation_per_particle_speed_min = speedMin; p.rotation_per_particle_speed_max = speedMax;
    apply(p);
}
float ParticleEmitter::getRotationPerParticleSpeedMin() const { return params().rotation_per_particle_speed_min; }
float ParticleEmitter::getRotationPerParticleSpeedMax() const { return params().rotation_per_particle_speed_max; }

void ParticleEmitter::setRotation(float speedMin, float speedMax, const Vector3& axis, const Vector3& axisVariance)
{
    t3d_emitter_params p = params();
    p.rotation_speed_min = speedMin; p.rotation_speed_max = speedMax;
    copy3(p.rotation_axis, axis); copy3(p.rotation_axis_var, axisVariance);
    apply(p);
}
float ParticleEmitter::getRotationSpeedMin() const { return params().rotation_speed_min; }
float ParticleEmitter::getRotationSpeedMax() const { return params().rotation_speed_max; }
Vector3 ParticleEmitter::getRotationAxis() const { return vec3(params().rotation_axis); }
Vector3 ParticleEmitter::getRotationAxisVariance() const { return vec3(params().rotation_axis_var); }

void ParticleEmitter::setSpriteAnimated(bool v) { t3d_emitter_params p = params(); p.sprite_animated = v; apply(p); }
bool ParticleEmitter::isSpriteAnimated() const { return params().sprite_animated != 0; }
void ParticleEmitter::setSpriteLooped(bool v) { t3d_emitter_params p = params(); p.sprite_looped = v; apply(p); }
bool ParticleEmitter::isSpriteLooped() const { return params().sprite_looped != 0; }
void ParticleEmitter::setSpriteFrameRandomOffset(int maxOffset)
{
    t3d_emitter_params p = params(); p.sprite_frame_random_offset = (uint32_t)maxOffset; apply(p);
}
int ParticleEmitter::getSpriteFrameRandomOffset() const { return (int)params().sprite_frame_random_offset; }
void ParticleEmitter::setSpriteFrameDuration(long duration) // PE.cpp:491-495
{
    t3d_emitter_params p = params(); p.sprite_frame_duration = duration; apply(p);
}
long ParticleEmitter::getSpriteFrameDuration() const { return (long)params().sprite_frame_duration; }

unsigned int ParticleEmitter::getSpriteWidth() const // PE.cpp:498-502
{
    float c[4];
    uint32_t n = 0;
    t3d_emitter_get_sprite_tex_coords(_h, &n, c, 1);
    return (unsigned int)std::fabs(params().texture_width * (c[2] - c[0]));
}
unsigned int ParticleEmitter::getSpriteHeight() const // PE.cpp:505-509
{
    float c[4];
    uint32_t n = 0;
    t3d_emitter_get_sprite_tex_coords(_h, &n, c, 1);
    return (unsigned int)std::fabs(params().texture_height * (c[3] - c[1]));
}

void ParticleEmitter::setSpriteTexCoords(unsigned int frameCount, float* texCoords) // PE.cpp:512-523
{
    if (t3d_emitter_set_sprite_tex_coords(_h, frameCount, texCoords) != T3D_OK) GP_ERROR("%s", t3d_last_error());
}
void ParticleEmitter::setSpriteFrameCoords(unsigned int frameCount, Rectangle* frameCoords) // PE.cpp:526-547
{
    static_assert(sizeof(Rectangle) == 4 * sizeof(float), "Rectangle is {x, y, width, height}");
    if (t3d_emitter_set_sprite_frame_rects(_h, frameCount, &frameCoords->x) != T3D_OK) GP_ERROR("%s", t3d_last_error());
}
void ParticleEmitter::setSpriteFrameCoords(unsigned int frameCount, int width, int height) // PE.cpp:550-582
{
    if (t3d_emitter_set_sprite_frame_coords(_h, frameCount, width, height) != T3D_OK) GP_ERROR("%s", t3d_last_error());
}
unsigned int ParticleEmitter::getSpriteFrameCount() const
{
    uint32_t n = 0;
    t3d_emitter_get_sprite_tex_coords(_h, &n, nullptr, 0);
    return n;
}

void ParticleEmitter::setOrbit(bool op, bool ov, bool oa) // PE.cpp:585-590
{
    t3d_emitter_params p = params();
    p.orbit_position = op; p.orbit_velocity = ov; p.orbit_acceleration = oa;
    apply(p);
}
bool ParticleEmitter::getOrbitPosition() const { return params().orbit_position != 0; }
bool ParticleEmitter::getOrbitVelocity() const { return params().orbit_velocity != 0; }
bool ParticleEmitter::getOrbitAcceleration() const { return params().orbit_acceleration != 0; }

void ParticleEmitter::setBlendMode(BlendMode blendMode) // PE.cpp:451-482
{
    if (blendMode < BLEND_NONE || blendMode > BLEND_MULTIPLIED)
    {
        GP_ERROR("Unsupported blend mode (%d).", (int)blendMode);
        return;
    }
    t3d_emitter_params p = params(); p.blend_mode = (int32_t)blendMode; apply(p);
}
ParticleEmitter::BlendMode ParticleEmitter::getBlendMode() const { return (BlendMode)params().blend_mode; }

ParticleEmitter::BlendMode ParticleEmitter::getBlendModeFromString(const std::string& s) // PE.cpp:682-710
{
    if (s == "BLEND_NONE" || s == "NONE" || s == "BLEND_OPAQUE" || s == "OPAQUE") return BLEND_NONE;
    if (s == "BLEND_ADDITIVE" || s == "ADDITIVE") return BLEND_ADDITIVE;
    if (s == "BLEND_MULTIPLIED" || s == "MULTIPLIED") return BLEND_MULTIPLIED;
    return BLEND_ALPHA;
}

void ParticleEmitter::update(float elapsedTime) // PE.cpp:713-832
{
    pushTransforms();
    if (t3d_emitter_update(_h, elapsedTime) != T3D_OK) GP_ERROR("update: %s", t3d_last_error());
}

unsigned int ParticleEmitter::draw(bool) // PE.cpp:835-888
{
    pushTransforms();
    uint32_t calls = 0;
    if (t3d_emitter_draw(_h, &calls) != T3D_OK) GP_ERROR("draw: %s", t3d_last_error());
    return calls;
}

Drawable* ParticleEmitter::clone(NodeCloneContext&) // PE.cpp:891-932
{
    t3d_emitter* h = nullptr;
    if (t3d_emitter_clone(_h, &h) != T3D_OK)
    {
        GP_ERROR("clone: %s", t3d_last_error());
        return nullptr;
    }
    ParticleEmitter* e = new ParticleEmitter();
    e->_h = h;
    if (_texture) // PE.cpp:894: the clone shares the texture
    {
        _texture->addRef();
        e->_texture = _texture;
    }
    return e;
}

void ParticleEmitter::setRandomSource(int mode, uint64_t seed)
{
    if (t3d_emitter_set_rng(_h, mode, seed) != T3D_OK) GP_ERROR("%s", t3d_last_error());
}
void ParticleEmitter::feedDraws(const int32_t* draws, size_t n) { t3d_emitter_feed_draws(_h, draws, n); }

const t3d_sprite_vertex* ParticleEmitter::getVertexStream(unsigned int* quadCount) const
{
    const t3d_sprite_vertex* p = nullptr;
    uint32_t n = 0;
    t3d_emitter_vertices(_h, &p, &n);
    if (quadCount) *quadCount = n;
    return p;
}

} // namespace tractor
