// The constants and the validity predicate of visual::VisualMapPoint that the depth path uses
// (le_vins/visual/visual_mappoint.h:38-51, visual_mappoint.cc:90-92).  The object graph itself stays in the estimator.
#ifndef LEVINS_HOST_VISUAL_MAPPOINT_H
#define LEVINS_HOST_VISUAL_MAPPOINT_H

namespace visual {

enum MapPointType {
    MAPPOINT_NONE = -1,
    MAPPOINT_TRIANGULATED = 0,
    MAPPOINT_DEPTH_ASSOCIATED = 1,
};

class VisualMapPoint {
public:
    static constexpr double DEFAULT_DEPTH = 10.0;
    static constexpr double NEAREST_DEPTH = 1;
    static constexpr double FARTHEST_DEPTH = 200;
    static bool isGoodDepth(double depth, double scale) { return ((depth > NEAREST_DEPTH) && (depth < FARTHEST_DEPTH * scale)); }
};

} // namespace visual
#endif
