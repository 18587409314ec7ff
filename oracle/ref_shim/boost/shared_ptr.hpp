// TEST INFRASTRUCTURE: boost::shared_ptr as the reference headers use it (see ../shim_eigen.h)
#pragma once
#include <memory>
namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
}
