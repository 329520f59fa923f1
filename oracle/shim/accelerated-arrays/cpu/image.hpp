#pragma once
#include "../image.hpp"
namespace accelerated { namespace cpu { struct Image {
    static std::unique_ptr<accelerated::Image::Factory> createFactory() { return std::unique_ptr<accelerated::Image::Factory>(new accelerated::Image::Factory()); }
}; } }
