// stand-in: ORBMatcher only passes a Map* through (ORBMatcher.cpp:524-526)
#pragma once
namespace mono_orb_slam3 { class Map {}; }
