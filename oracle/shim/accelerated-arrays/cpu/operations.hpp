#pragma once
#include "../standard_ops.hpp"
namespace accelerated { namespace cpu { namespace operations {
inline std::unique_ptr<accelerated::operations::StandardFactory> createFactory(Processor &) {
    return std::unique_ptr<accelerated::operations::StandardFactory>(new accelerated::operations::StandardFactory());
}
} } }
