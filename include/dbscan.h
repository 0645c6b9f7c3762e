// dbscan.h — the clustering step of the reference's evaluator (reference dogm/demo/utils/include/dbscan.h:9-48,
// dogm/demo/utils/dbscan.cpp:19-114) for users of the drop-in headers: same types and the same partition of the points,
// header only.  The reference's procedure is kept behaviour for behaviour, because the evaluator's error figures depend
// on which cells end up in which cluster:
//   * a point is identified by its coordinates (Point::operator==);
//   * after the first region query of a new cluster the reference labels the first |neighbours| points OF THE INPUT
//     ARRAY, not the neighbours themselves (dbscan.cpp:66-69); the neighbours get their label when a later region query
//     finds them again;
//   * neighbours still unclassified when found again are queued a second time; all queue entries equal to the point just
//     processed are dropped together (dbscan.cpp:71,95).
// Implemented on indices instead of copied points; O(n^2) like the reference (a few hundred dynamic cells per scan).
#pragma once

#include <cmath>
#include <cstddef>
#include <vector>

template <typename T>
struct Point
{
    float x, y;
    T data;
    int cluster_id;

    bool operator==(const Point<T>& other) const { return x == other.x && y == other.y; }
};

constexpr int UNCLASSIFIED = -2;
constexpr int NOISE = -1;

template <typename T>
using Cluster = std::vector<Point<T>>;

template <typename T>
using Clusters = std::vector<Cluster<T>>;

template <typename T>
class DBSCAN
{
  public:
    DBSCAN(float eps, int min_cells) : eps(eps), min_cells(min_cells) {}

    Clusters<T> cluster(const std::vector<Point<T>>& points) const
    {
        std::vector<Point<T>> labelled = points;
        int next_id = NOISE + 1;
        for (std::size_t i = 0; i < labelled.size(); i++)
            if (labelled[i].cluster_id == UNCLASSIFIED && grow(labelled, i, next_id))
                next_id++;
        Clusters<T> clusters(static_cast<std::size_t>(next_id), Cluster<T>());
        for (const Point<T>& p : labelled)
            if (p.cluster_id >= 0 && p.cluster_id < next_id)
                clusters[static_cast<std::size_t>(p.cluster_id)].push_back(p);
        return clusters;
    }

  private:
    std::vector<std::size_t> neighbours(const std::vector<Point<T>>& pts, std::size_t q) const
    {
        std::vector<std::size_t> found;
        for (std::size_t i = 0; i < pts.size(); i++)
            if (sqrtf(powf(pts[q].x - pts[i].x, 2) + powf(pts[q].y - pts[i].y, 2)) <= eps)
                found.push_back(i);
        return found;
    }

    // queue entries that denote the same point (same coordinates) as pts[who] leave the queue
    static void drop(std::vector<std::size_t>& queue, const std::vector<Point<T>>& pts, std::size_t who)
    {
        std::size_t kept = 0;
        for (std::size_t k = 0; k < queue.size(); k++)
            if (!(pts[queue[k]] == pts[who]))
                queue[kept++] = queue[k];
        queue.resize(kept);
    }

    bool grow(std::vector<Point<T>>& pts, std::size_t start, int id) const
    {
        std::vector<std::size_t> queue = neighbours(pts, start);
        if (queue.size() < static_cast<std::size_t>(min_cells))
        {
            pts[start].cluster_id = NOISE;
            return false;
        }
        for (std::size_t k = 0; k < queue.size(); k++) // dbscan.cpp:66-69 (see the header comment)
            pts[k].cluster_id = id;
        drop(queue, pts, start);
        while (!queue.empty())
        {
            const std::size_t current = queue.front();
            const std::vector<std::size_t> around = neighbours(pts, current);
            if (around.size() >= static_cast<std::size_t>(min_cells))
            {
                for (std::size_t r : around)
                {
                    const int label = pts[r].cluster_id;
                    if (label != UNCLASSIFIED && label != NOISE)
                        continue;
                    if (label == UNCLASSIFIED)
                        queue.push_back(r);
                    std::size_t first = 0; // the first point with these coordinates takes the label
                    while (!(pts[first] == pts[r]))
                        first++;
                    pts[first].cluster_id = id;
                }
            }
            drop(queue, pts, current);
        }
        return true;
    }

    float eps;
    int min_cells;
};
