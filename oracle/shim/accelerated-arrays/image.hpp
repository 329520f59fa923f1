// stand-in for accelerated-arrays (absent third-party library): a CPU image header over an 8-bit plane.  The GPU
// (OpenGL) paths of the reference compile against these declarations and are never executed in oracle/_ref.
#pragma once
#include <cstdlib>
#include <memory>
#include "future.hpp"
namespace accelerated {
struct Processor {
    virtual ~Processor() = default;
    static std::unique_ptr<Processor> createInstant() { return std::unique_ptr<Processor>(new Processor()); }
};
struct Image {
    enum class StorageType { CPU, GPU_OPENGL };
    enum class DataType { UINT8 };
    enum class Interpolation { NEAREST, LINEAR };
    enum class Border { ZERO, MIRROR, CLAMP };
    int width = 0, height = 0, channels = 1;
    DataType dataType = DataType::UINT8;
    StorageType storageType = StorageType::CPU;
    unsigned char *data = nullptr;   // shim: CPU plane
    size_t stride = 0;
    virtual ~Image() = default;
    struct Factory {
        virtual ~Factory() = default;
        virtual std::unique_ptr<Image> create(int, int, int, DataType) { std::abort(); }
        virtual std::unique_ptr<Image> createLike(const Image &) { std::abort(); }
    };
};
}  // namespace accelerated
