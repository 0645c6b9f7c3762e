#pragma once
#include "mat4x4.hpp"
#include "vec4.hpp"
