// metrics.h — error statistics of the reference's evaluator (reference dogm/demo/utils/include/metrics.h:11-63,
// dogm/demo/utils/metrics.cpp:8-56): mean absolute error and root mean squared error of cluster position [m] and velocity
// [m/s] against the simulated vehicle it was matched with.  Header only.
#pragma once

#include "simulator.h"
#include "types.h"

#include <cmath>

class Metric
{
  public:
    Metric() : cumulative_error{}, number_of_detections(0) {}
    virtual ~Metric() = default;

    virtual void reset()
    {
        cumulative_error = {};
        number_of_detections = 0;
    }
    virtual PointWithVelocity addObjectDetection(const PointWithVelocity& cluster_mean, const Vehicle& vehicle) = 0;
    virtual PointWithVelocity computeErrorStatistic() = 0;
    int numberOfDetections() const { return number_of_detections; }

  protected:
    static PointWithVelocity computeError(const PointWithVelocity& cluster_mean, const Vehicle& vehicle)
    {
        PointWithVelocity error{};
        error.x += cluster_mean.x - vehicle.pos[0];
        error.y += cluster_mean.y - vehicle.pos[1];
        error.v_x += cluster_mean.v_x - vehicle.vel[0];
        error.v_y += cluster_mean.v_y - vehicle.vel[1];
        return error;
    }

    template <typename F>
    PointWithVelocity accumulate(const PointWithVelocity& cluster_mean, const Vehicle& vehicle, F&& term)
    {
        const PointWithVelocity error = computeError(cluster_mean, vehicle);
        cumulative_error.x += term(error.x);
        cumulative_error.y += term(error.y);
        cumulative_error.v_x += term(error.v_x);
        cumulative_error.v_y += term(error.v_y);
        ++number_of_detections;
        return error;
    }

    template <typename F>
    PointWithVelocity statistic(F&& finish) const
    {
        PointWithVelocity out;
        out.x = finish(cumulative_error.x / number_of_detections);
        out.y = finish(cumulative_error.y / number_of_detections);
        out.v_x = finish(cumulative_error.v_x / number_of_detections);
        out.v_y = finish(cumulative_error.v_y / number_of_detections);
        return out;
    }

    PointWithVelocity cumulative_error;
    int number_of_detections;
};

class MAE : public Metric // metrics.cpp:8-32
{
  public:
    PointWithVelocity addObjectDetection(const PointWithVelocity& cluster_mean, const Vehicle& vehicle) override
    {
        return accumulate(cluster_mean, vehicle, [](float e) { return std::fabs(e); });
    }
    PointWithVelocity computeErrorStatistic() override
    {
        return statistic([](float mean) { return mean; });
    }
};

class RMSE : public Metric // metrics.cpp:34-56
{
  public:
    PointWithVelocity addObjectDetection(const PointWithVelocity& cluster_mean, const Vehicle& vehicle) override
    {
        return accumulate(cluster_mean, vehicle, [](float e) { return powf(e, 2.0f); });
    }
    PointWithVelocity computeErrorStatistic() override
    {
        return statistic([](float mean) { return sqrtf(mean); });
    }
};
