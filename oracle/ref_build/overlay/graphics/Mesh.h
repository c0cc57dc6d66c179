// Overlay stub for graphics/Mesh.h: the vertex-format description and primitive enum SpriteBatch::create
// hands to MeshBatch::create (SpriteBatch.cpp:136-148). No vertex buffers, no bounding volumes.
#pragma once
#include <vector>
#include "utils/Ref.h"
namespace tractor
{
class VertexFormat
{
  public:
    enum Usage { POSITION = 1, NORMAL, COLOR, TANGENT, BINORMAL, BLENDWEIGHTS, BLENDINDICES, TEXCOORD0 };
    struct Element
    {
        Usage usage{ POSITION };
        unsigned int size{ 0 };
        Element() = default;
        Element(Usage u, unsigned int s) : usage(u), size(s) {}
    };
    VertexFormat(const Element* elements, unsigned int count) : _elements(elements, elements + count) {}
    unsigned int getVertexSize() const { unsigned int n = 0; for (auto& e : _elements) n += e.size * sizeof(float); return n; }
  private:
    std::vector<Element> _elements;
};
class Mesh : public Ref
{
  public:
    enum PrimitiveType { TRIANGLES = 4, TRIANGLE_STRIP = 5, LINES = 1, LINE_STRIP = 3, POINTS = 0 };
};
} // namespace tractor
