// Overlay stub for scene/Scene.h: only the active camera's node is consulted (ParticleEmitter.cpp:858-861).
#pragma once
#include "scene/Node.h"
namespace tractor
{
class Camera
{
  public:
    Node* getNode() const { return _node; }
    Node* _node{ nullptr };
};
class Scene
{
  public:
    Camera* getActiveCamera() const { return _camera; }
    Camera* _camera{ nullptr };
};
} // namespace tractor
