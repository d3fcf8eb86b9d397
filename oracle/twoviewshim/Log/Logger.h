// stand-in for the reference's Log/Logger.h: TwoViewReconstruction.cpp streams progress lines into initial_logger
#pragma once
#include <string>
namespace mono_orb_slam3 {
    struct Logger { template <class T> Logger &operator<<(const T &) { return *this; } };
    static const std::string titles[3] = {"", "", ""};
    static Logger initial_logger;
}
