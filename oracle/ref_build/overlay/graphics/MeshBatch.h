// Overlay stub for graphics/MeshBatch.h: a capture sink. add() keeps the exact bytes SpriteBatch hands
// over (the reference memcpy's them into a client-side array, MeshBatch.cpp:87-148) so the harness can
// read back the vertex stream of one draw(); start() clears, finish() is a no-op and draw() (GL dropped) only
// notes that the batch was submitted.
#pragma once
#include <cstring>
#include <vector>
#include "graphics/Mesh.h"
#include "renderer/Material.h"
namespace tractor
{
class MeshBatch
{
  public:
    static MeshBatch* create(const VertexFormat& vertexFormat, Mesh::PrimitiveType, Material* material,
                             bool /*indexed*/, unsigned int initialCapacity = 1024, unsigned int = 1024)
    {
        MeshBatch* b = new MeshBatch();
        b->_vertexSize = vertexFormat.getVertexSize();
        b->_material = material;
        material->addRef();
        b->_bytes.reserve((size_t)initialCapacity * 4 * b->_vertexSize);
        return b;
    }
    ~MeshBatch() { if (_material) _material->release(); }
    template <class T> void add(const T* vertices, unsigned int vertexCount, unsigned short* /*indices*/ = nullptr,
                                unsigned int /*indexCount*/ = 0)
    {
        if (discard) { _vertexCount += vertexCount; return; }
        const unsigned char* p = reinterpret_cast<const unsigned char*>(vertices);
        _bytes.insert(_bytes.end(), p, p + (size_t)vertexCount * sizeof(T));
        _vertexCount += vertexCount;
    }
    void start() { _bytes.clear(); _vertexCount = 0; _started = true; }
    bool isStarted() const { return _started; }
    void finish() { _started = false; }
    void draw() { submitted = true; } // SpriteBatch::finish() -> MeshBatch::draw(): this is where GL would be called
    Material* getMaterial() const { return _material; }
    const std::vector<unsigned char>& capturedBytes() const { return _bytes; }
    size_t capturedVertexCount() const { return _vertexCount; }
    bool discard{ false };
    bool submitted{ false }; // set when the batch was handed to "GL" since the harness last cleared it
  private:
    Material* _material{ nullptr };
    unsigned int _vertexSize{ 0 };
    std::vector<unsigned char> _bytes;
    size_t _vertexCount{ 0 };
    bool _started{ false };
};
} // namespace tractor
