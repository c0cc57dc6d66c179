// Overlay stub for renderer/Texture.h: the particle path needs only width/height (ParticleEmitter.cpp:244-247)
// and a Sampler object to hang on the SpriteBatch. Texture::create is defined in ref_capi.cpp and reads
// the PNG IHDR for the dimensions (no decoding, no GL).
#pragma once
#include <string>
#include "utils/Ref.h"
namespace tractor
{
class Texture : public Ref
{
  public:
    enum Filter { NEAREST, LINEAR, NEAREST_MIPMAP_NEAREST, LINEAR_MIPMAP_NEAREST, NEAREST_MIPMAP_LINEAR, LINEAR_MIPMAP_LINEAR };
    enum Type { TEXTURE_2D, TEXTURE_CUBE };
    enum Wrap { REPEAT, CLAMP };
    static Texture* create(const std::string& path, bool generateMipmaps = false);
    unsigned int getWidth() const { return _width; }
    unsigned int getHeight() const { return _height; }
    Type getType() const { return TEXTURE_2D; }
    unsigned int _width{ 0 };
    unsigned int _height{ 0 };

    class Sampler : public Ref
    {
      public:
        static Sampler* create(Texture* texture) { Sampler* s = new Sampler(); s->_texture = texture; texture->addRef(); return s; }
        ~Sampler() { if (_texture) _texture->release(); }
        void setFilterMode(Filter, Filter) {}
        void setWrapMode(Wrap, Wrap, Wrap = REPEAT) {}
        Texture* getTexture() const { return _texture; }
      private:
        Texture* _texture{ nullptr };
    };
};
} // namespace tractor
