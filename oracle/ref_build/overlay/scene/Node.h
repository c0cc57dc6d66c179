// Overlay stub for scene/Node.h. The particle path reads three things from its node:
// the world matrix (ParticleEmitter.cpp:299), the view-projection matrix (:848) and, through
// getScene()->getActiveCamera()->getNode(), the camera node's world matrix (:860-861).
#pragma once
#include "graphics/Drawable.h"
#include "math/Matrix.h"
namespace tractor
{
class Scene;
class NodeCloneContext {};
class Node
{
  public:
    const Matrix& getWorldMatrix() const { return _world; }
    const Matrix& getViewProjectionMatrix() const { return _viewProjection; }
    Scene* getScene() const { return _scene; }
    void setDrawable(Drawable* drawable);   // defined in ref_capi.cpp (Drawable befriends Node)
    Matrix _world;
    Matrix _viewProjection;
    Scene* _scene{ nullptr };
    Drawable* _drawable{ nullptr };
};
} // namespace tractor
