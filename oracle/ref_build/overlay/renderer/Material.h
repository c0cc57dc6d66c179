// Overlay stub for renderer/Material.h: SpriteBatch::create wraps its effect in a Material and binds
// two parameters (SpriteBatch.cpp:124-171). Parameters accept and ignore their values.
#pragma once
#include <cfloat> // (Font.cpp uses FLT_MAX and gets it through the real Material.h)
#include <string>
#include "graphics/Effect.h"
#include "renderer/RenderState.h"
namespace tractor
{
class MaterialParameter
{
  public:
    template <class T> void setValue(T) {}
    template <class T> void setVector2(T) {} // (Font.cpp:367: the distance-field cutoff)
    template <class C, class R> void bindValue(C*, R (C::*)() const) {}
};
class Material : public RenderState
{
  public:
    static Material* create(Effect* effect) { Material* m = new Material(); m->_effect = effect; if (effect) effect->addRef(); return m; } // shares the effect, as the engine's does
    ~Material() { if (_effect) _effect->release(); }
    MaterialParameter* getParameter(const std::string&) { return &_param; }
  private:
    Effect* _effect{ nullptr };
    MaterialParameter _param;
};
} // namespace tractor
