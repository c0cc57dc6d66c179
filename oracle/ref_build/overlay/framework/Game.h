// Overlay stub shadowing the reference's framework/Game.h (which drags in audio, physics, input).
// SpriteBatch.cpp only asks the Game singleton for a viewport to seed a default ortho matrix
// (reference src/graphics/SpriteBatch.cpp:162-169); ParticleEmitter::draw overrides it anyway.
#pragma once
#include "graphics/Rectangle.h"
namespace tractor
{
class Game
{
  public:
    static Game* getInstance() { static Game g; return &g; }
    const Rectangle& getViewport() const { return _viewport; }
  private:
    Rectangle _viewport{ 0.0f, 0.0f, 1280.0f, 720.0f };
};
} // namespace tractor
