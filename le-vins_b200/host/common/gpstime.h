// common/gpstime.h of the reference (le_vins/common/gpstime.h:25-42): same class, same arithmetic.
#ifndef LEVINS_HOST_GPSTIME_H
#define LEVINS_HOST_GPSTIME_H

#include <cmath>

#define GPS_LEAP_SECOND 18

class GpsTime {
public:
    static void gps2unix(int week, double sow, double &unixs) { unixs = sow + week * 604800 + 315964800 - GPS_LEAP_SECOND; }

    static void unix2gps(double unixs, int &week, double &sow) {
        double seconds = unixs + GPS_LEAP_SECOND - 315964800;

        week = floor(seconds / 604800);
        sow  = seconds - week * 604800;
    }
};

#endif
