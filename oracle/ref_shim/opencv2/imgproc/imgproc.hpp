// ORACLE build shim (test infrastructure): the reference includes the per-module OpenCV headers;
// every declaration it needs lives in opencv2/opencv.hpp of this shim.
#include "opencv2/opencv.hpp"
