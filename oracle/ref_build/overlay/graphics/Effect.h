// Overlay stub for graphics/Effect.h: a shader program with exactly one sampler uniform, which is
// what SpriteBatch::create looks for (SpriteBatch.cpp:105-121). No GLSL is compiled.
#pragma once
#include <string>
#include "math/Matrix.h"
#include "math/Vector2.h"
#include "math/Vector3.h"
#include "math/Vector4.h"
#include "renderer/Texture.h"
#include "utils/Ref.h"
namespace tractor
{
class Uniform
{
  public:
    unsigned int getType() const { return 0x8B5E; /* GL_SAMPLER_2D */ }
    const std::string& getName() const { return _name; }
  private:
    std::string _name{ "u_texture" };
};
class Effect : public Ref
{
  public:
    static Effect* createFromFile(const std::string&, const std::string&, const char* = nullptr) { return new Effect(); }
    static Effect* createFromFile(const std::string&, const std::string&, const std::string&) { return new Effect(); } // (Font.cpp:122)
    unsigned int getUniformCount() const { return 1; }
    Uniform* getUniform(unsigned int) { return &_uniform; }
  private:
    Uniform _uniform;
};
} // namespace tractor
