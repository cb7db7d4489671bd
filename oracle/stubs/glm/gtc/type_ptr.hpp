// TEST INFRASTRUCTURE. (see ../../glad/glad.h)
#pragma once
#include "../mat4x4.hpp"
namespace glm { template <typename T> auto value_ptr(T& t) { return &t.v[0]; } }
