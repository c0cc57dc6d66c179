// Overlay stub for renderer/RenderState.h: blend enums and a StateBlock that records what
// ParticleEmitter::setBlendMode asks for (ParticleEmitter.cpp:451-482). No GL.
#pragma once
#include "utils/Ref.h"
namespace tractor
{
class RenderState : public Ref
{
  public:
    enum Blend { BLEND_ZERO, BLEND_ONE, BLEND_SRC_COLOR, BLEND_ONE_MINUS_SRC_COLOR, BLEND_DST_COLOR,
                 BLEND_ONE_MINUS_DST_COLOR, BLEND_SRC_ALPHA, BLEND_ONE_MINUS_SRC_ALPHA, BLEND_DST_ALPHA,
                 BLEND_ONE_MINUS_DST_ALPHA, BLEND_CONSTANT_ALPHA, BLEND_ONE_MINUS_CONSTANT_ALPHA, BLEND_SRC_ALPHA_SATURATE };
    class StateBlock
    {
      public:
        void setBlend(bool b) { blend = b; }
        void setBlendSrc(Blend b) { src = b; }
        void setBlendDst(Blend b) { dst = b; }
        void setDepthWrite(bool b) { depthWrite = b; }
        void setDepthTest(bool b) { depthTest = b; }
        bool blend{ false }, depthWrite{ true }, depthTest{ false };
        Blend src{ BLEND_ONE }, dst{ BLEND_ZERO };
    };
    StateBlock* getStateBlock() { return &_state; }
  protected:
    StateBlock _state;
};
} // namespace tractor
