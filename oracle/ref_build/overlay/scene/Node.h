// Overlay stub for scene/Node.h. The particle path reads three things from its node:
// the world matrix (ParticleEmitter.cpp:299), the view-projection matrix (:848) and, through
// getScene()->getActiveCamera()->getNode(), the camera node's world matrix (:860-861).
#pragma once
#include "graphics/Drawable.h"
#include "math/Matrix.h"
#include "math/Vector3.h"
namespace tractor
{
class Scene;
class NodeCloneContext {};
class Node
{
  public:
    const Matrix& getWorldMatrix() const { return _world; }
    const Matrix& getViewProjectionMatrix() const { return _viewProjection; }
    // TileSet::draw (src/graphics/TileSet.cpp:182-205) also asks for these two
    const Matrix& getProjectionMatrix() const { return _projection; }
    Vector3 getTranslationWorld() const { return Vector3(_world.m[12], _world.m[13], _world.m[14]); }
    Scene* getScene() const { return _scene; }
    void setDrawable(Drawable* drawable);   // defined in ref_capi.cpp (Drawable befriends Node)
    Matrix _world;
    Matrix _viewProjection;
    Matrix _projection;
    Scene* _scene{ nullptr };
    Drawable* _drawable{ nullptr };
};
} // namespace tractor
