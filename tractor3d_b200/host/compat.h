// compat.h -- the handful of engine types the ParticleEmitter interface mentions, for building the facade
// outside the engine.  Inside Tractor3D define T3D_WITH_ENGINE_HEADERS and the engine's own
// math/Vector*.h, math/Matrix.h, graphics/Rectangle.h, utils/Ref.h, graphics/Drawable.h, scene/Node.h are
// used instead (same member names: x/y/z/w, m[16], getWorldMatrix(), addRef()/release()).
#pragma once

#ifndef T3D_WITH_ENGINE_HEADERS
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace tractor
{

struct Vector2 { float x{ 0 }, y{ 0 }; Vector2() = default; Vector2(float x_, float y_) : x(x_), y(y_) {} };
struct Vector3
{
    float x{ 0 }, y{ 0 }, z{ 0 };
    Vector3() = default;
    Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    void set(float x_, float y_, float z_) { x = x_; y = y_; z = z_; }
    static Vector3 zero() { return Vector3(); }
    static Vector3 one() { return Vector3(1, 1, 1); }
};
struct Vector4
{
    float x{ 0 }, y{ 0 }, z{ 0 }, w{ 0 };
    Vector4() = default;
    Vector4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    void set(float x_, float y_, float z_, float w_) { x = x_; y = y_; z = z_; w = w_; }
    static Vector4 zero() { return Vector4(); }
    static Vector4 one() { return Vector4(1, 1, 1, 1); }
};
struct Rectangle
{
    float x{ 0 }, y{ 0 }, width{ 0 }, height{ 0 };
    Rectangle() = default;
    Rectangle(float w, float h) : width(w), height(h) {}
    Rectangle(float x_, float y_, float w, float h) : x(x_), y(y_), width(w), height(h) {}
};
struct Matrix
{
    float m[16]; // column-major, reference include/math/Matrix.h:62
    Matrix() { std::memset(m, 0, sizeof(m)); m[0] = m[5] = m[10] = m[15] = 1.0f; }
    explicit Matrix(const float* src) { std::memcpy(m, src, sizeof(m)); }
    static const Matrix& identity() { static Matrix i; return i; }
};

// Intrusive reference count, reference src/utils/Ref.cpp:28-61.
class Ref
{
  public:
    void addRef() { ++_refCount; }
    void release() { if (--_refCount <= 0) delete this; }
    unsigned int getRefCount() const { return _refCount; }
  protected:
    Ref() = default;
    virtual ~Ref() = default;
  private:
    int _refCount{ 1 };
};

class Node;
class NodeCloneContext {};

// reference include/graphics/Drawable.h
class Drawable
{
    friend class Node;
  public:
    virtual ~Drawable() = default;
    virtual unsigned int draw(bool wireframe = false) = 0;
    Node* getNode() const noexcept { return _node; }
  protected:
    virtual Drawable* clone(NodeCloneContext& context) = 0;
    virtual void setNode(Node* node) { _node = node; }
    Node* _node{ nullptr };
};

// The three things ParticleEmitter asks of its node (PE.cpp:299, 848, 858-861).
class Camera;
class Scene
{
  public:
    Camera* getActiveCamera() const { return _camera; }
    void setActiveCamera(Camera* c) { _camera = c; }
  private:
    Camera* _camera{ nullptr };
};
class Node
{
  public:
    const Matrix& getWorldMatrix() const { return _world; }
    const Matrix& getViewProjectionMatrix() const { return _viewProjection; }
    Scene* getScene() const { return _scene; }
    void setWorldMatrix(const Matrix& m) { _world = m; }
    void setScene(Scene* s) { _scene = s; }
    void setDrawable(Drawable* d) // reference src/scene/Node.cpp:746-767
    {
        if (_drawable) _drawable->setNode(nullptr);
        _drawable = d;
        if (d) d->setNode(this);
    }
  private:
    Matrix _world, _viewProjection;
    Scene* _scene{ nullptr };
    Drawable* _drawable{ nullptr };
};
class Camera
{
  public:
    Node* getNode() const { return _node; }
    void setNode(Node* n) { _node = n; }
  private:
    Node* _node{ nullptr };
};

} // namespace tractor

// reference include/pch.h:83-103.  The reference's GP_ERROR logs, breaks and exits unless
// GP_ERRORS_AS_WARNINGS is defined; that switch is honoured here.
#define GP_WARN(...) do { std::fprintf(stderr, "%s -- ", __FUNCTION__); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#ifdef GP_ERRORS_AS_WARNINGS
#define GP_ERROR GP_WARN
#else
#define GP_ERROR(...) do { GP_WARN(__VA_ARGS__); std::exit(-1); } while (0)
#endif
#endif // T3D_WITH_ENGINE_HEADERS
