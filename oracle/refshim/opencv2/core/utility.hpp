// TEST INFRASTRUCTURE ONLY - forwards to the one-header OpenCV stand-in (oracle/refshim/opencv2/core.hpp).
#pragma once
#include <opencv2/core.hpp>
