// TEST INFRASTRUCTURE ONLY — stand-in for <boost/serialization/serialization.hpp> (not installed here): DBoW2's BowVector.h /
// FeatureVector.h only befriend `access` and name `base_object` inside a serialize() template that oracle/_ref never instantiates.
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return d; }
} }
