// Overlay stub for scene/Node.h. The particle path reads three things from its node:
// the world matrix (ParticleEmitter.cpp:299), the view-projection matrix (:848) and, through
// getScene()->getActiveCamera()->getNode(), the camera node's world matrix (:860-861).
#pragma once
#include "graphics/Drawable.h"
#include "math/Matrix.h"
#include "math/Quaternion.h"
#include "math/Vector3.h"
namespace tractor
{
class Scene;
class Node;
class NodeCloneContext
{
  public:
    Node* findClonedNode(const Node*) { return nullptr; } // (Sprite::clone, src/ui/Sprite.cpp:587; never called by the harness)
};
class Node
{
  public:
    const Matrix& getWorldMatrix() const { return _world; }
    const Matrix& getViewProjectionMatrix() const { return _viewProjection; }
    // TileSet::draw (src/graphics/TileSet.cpp:182-205) also asks for these two
    const Matrix& getProjectionMatrix() const { return _projection; }
    Vector3 getTranslationWorld() const { return Vector3(_world.m[12], _world.m[13], _world.m[14]); }
    // ui/Sprite::draw (src/ui/Sprite.cpp:495-574) also asks for the node's rotation and scale
    const Quaternion& getRotation() const { return _rotation; }
    float getScaleX() const { return _scaleX; }
    float getScaleY() const { return _scaleY; }
    Scene* getScene() const { return _scene; }
    void setDrawable(Drawable* drawable);   // defined in ref_capi.cpp (Drawable befriends Node)
    Matrix _world;
    Matrix _viewProjection;
    Matrix _projection;
    Quaternion _rotation;
    float _scaleX{ 1.0f }, _scaleY{ 1.0f };
    Scene* _scene{ nullptr };
    Drawable* _drawable{ nullptr };
};
} // namespace tractor
