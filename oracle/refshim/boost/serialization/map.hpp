// TEST INFRASTRUCTURE ONLY — stand-in for <boost/serialization/map.hpp>, see serialization.hpp
#pragma once
#include "serialization.hpp"
