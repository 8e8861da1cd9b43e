#pragma once
#include <pcl/pcl_config.h>
#include <memory>
namespace pcl {
template <typename T>
using shared_ptr = std::shared_ptr<T>;
}
