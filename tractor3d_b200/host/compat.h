// compat.h -- the handful of engine types the ParticleEmitter interface mentions, for building the facade
// outside the engine.  Inside Tractor3D define T3D_WITH_ENGINE_HEADERS and the engine's own headers are used
// instead (same member names: x/y/z/w, m[16], getWorldMatrix(), addRef()/release(), Properties' getters);
// tests/test_engine_headers.py builds the facade both ways.
#pragma once

#ifdef T3D_WITH_ENGINE_HEADERS
#include "pch.h"

#include "graphics/Drawable.h"
#include "graphics/Rectangle.h"
#include "math/Matrix.h"
#include "math/Vector2.h"
#include "math/Vector3.h"
#include "math/Vector4.h"
#include "renderer/Texture.h"
#include "scene/Node.h"
#include "scene/Properties.h"
#include "scene/Scene.h"
#include "utils/Ref.h"
#else
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/t3d_particles.h"

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

// reference include/renderer/Texture.h, as far as the particle path looks: it only ever asks a texture for its size
// (PE.cpp:244-247), so create() reads the PNG header and nothing else.
class Texture : public Ref
{
  public:
    static Texture* create(const std::string& path, bool generateMipmaps = false)
    {
        (void)generateMipmaps;
        uint32_t w = 0, h = 0;
        if (t3d_host_png_size(path.c_str(), &w, &h) != T3D_OK) return nullptr;
        Texture* t = new Texture();
        t->_path = path;
        t->_width = w;
        t->_height = h;
        return t;
    }
    // (no engine equivalent: a texture that only exists as a size, for callers that have no image file)
    static Texture* create(unsigned int width, unsigned int height)
    {
        Texture* t = new Texture();
        t->_width = width;
        t->_height = height;
        return t;
    }
    const std::string& getPath() const { return _path; }
    unsigned int getWidth() const { return _width; }
    unsigned int getHeight() const { return _height; }
  private:
    std::string _path;
    unsigned int _width{ 0 }, _height{ 0 };
};

// reference include/scene/Properties.h, the read side: namespaces of `key = value` strings with typed getters.  A view
// over the library's t3d_properties tree; the object returned by create() owns the tree, children belong to their parent.
struct Property
{
    std::string name, value;
};
class Properties
{
  public:
    static Properties* create(const std::string& url) // Properties.cpp:62-100
    {
        t3d_properties* root = nullptr;
        if (t3d_properties_load(url.c_str(), &root) != T3D_OK) return nullptr;
        Properties* p = new Properties(root);
        p->_owned = root;
        return p;
    }
    ~Properties() { if (_owned) t3d_properties_free(_owned); }
    Properties(const Properties&) = delete;
    Properties& operator=(const Properties&) = delete;

    const std::string& getNamespace() const noexcept { return _namespace; }
    const std::string& getId() const noexcept { return _id; }
    void rewind() { _nextNamespace = 0; _nextProperty = 0; }
    Properties* getNextNamespace()
    {
        if (_nextNamespace >= t3d_properties_namespace_count(_p)) return nullptr;
        const size_t i = _nextNamespace++;
        if (_children.size() <= i) _children.resize(i + 1);
        if (!_children[i]) _children[i].reset(new Properties(t3d_properties_get_namespace(_p, i)));
        return _children[i].get();
    }
    Property* getNextProperty()
    {
        if (_nextProperty >= t3d_properties_count(_p)) return nullptr;
        _current.name = t3d_properties_key(_p, _nextProperty);
        _current.value = t3d_properties_value(_p, _nextProperty);
        ++_nextProperty;
        return &_current;
    }
    bool exists(const std::string& name) const { return t3d_properties_get(_p, name.c_str()) != nullptr; }
    std::string getString(const std::string& name, const std::string& defaultValue = std::string()) const
    {
        const char* v = t3d_properties_get(_p, name.c_str());
        return v ? std::string(v) : defaultValue;
    }
    bool setString(const std::string& name, const std::string& value) { return t3d_properties_set(_p, name.c_str(), value.c_str()) == T3D_OK; }
    bool getBool(const std::string& name, bool defaultValue = false) const { return t3d_properties_get_bool(_p, name.c_str(), defaultValue) != 0; }
    int getInt(const std::string& name) const { return t3d_properties_get_int(_p, name.c_str()); }
    float getFloat(const std::string& name) const { return t3d_properties_get_float(_p, name.c_str()); }
    long getLong(const std::string& name) const { return (long)t3d_properties_get_long(_p, name.c_str()); }
    bool getVector3(const std::string& name, Vector3* out) const
    {
        float v[3];
        const bool ok = t3d_properties_get_vector(_p, name.c_str(), v, 3) != 0;
        if (out) out->set(v[0], v[1], v[2]);
        return ok;
    }
    bool getVector4(const std::string& name, Vector4* out) const
    {
        float v[4];
        const bool ok = t3d_properties_get_vector(_p, name.c_str(), v, 4) != 0;
        if (out) out->set(v[0], v[1], v[2], v[3]);
        return ok;
    }
    bool getPath(const std::string& name, std::string& path) const
    {
        char buf[1024];
        if (t3d_properties_get_path(_p, name.c_str(), buf, sizeof(buf)) != T3D_OK) return false;
        path = buf;
        return true;
    }
    t3d_properties* handle() const { return _p; }

  private:
    explicit Properties(t3d_properties* p) : _p(p), _namespace(t3d_properties_namespace(p)), _id(t3d_properties_id(p)) {}
    t3d_properties* _p{ nullptr };
    t3d_properties* _owned{ nullptr };
    std::string _namespace, _id;
    size_t _nextNamespace{ 0 }, _nextProperty{ 0 };
    Property _current;
    std::vector<std::unique_ptr<Properties>> _children;
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
