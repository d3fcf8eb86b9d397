// stand-in: ORBMatcher.cpp includes this header but its only use (lie::Hatf, ORBMatcher.cpp:431) is commented out
#pragma once
