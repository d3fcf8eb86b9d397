// Name-only stand-in for boost::serialization, so that the reference's vendored DBoW2 headers (BowVector.h, FeatureVector.h) compile
// VERBATIM in an image without boost.  TEST INFRASTRUCTURE ONLY.  The `serialize` member templates that mention these names are
// never instantiated by the oracle harness (nothing is archived).
#pragma once
namespace boost { namespace serialization {
    class access;
    template <class Base, class Derived> Base &base_object(Derived &d) { return static_cast<Base &>(d); }
}}
