// stand-in for the reference's Log/Logger.h: ORBMatcher.cpp:410-411 streams a statistics line into tracker_logger
#pragma once
#include <string>
namespace mono_orb_slam3 {
    struct Logger { template <class T> Logger &operator<<(const T &) { return *this; } };
    static const std::string titles[3] = {"", "", ""};
    static Logger tracker_logger;
}
