// types.h — reference dogm/demo/utils/include/types.h:8-14
#pragma once

struct PointWithVelocity
{
    float x{0.0f};
    float y{0.0f};
    float v_x{0.0f};
    float v_y{0.0f};
};
