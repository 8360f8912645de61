// ORACLE build shim (test infrastructure): see opencv2/opencv.hpp of this shim.
#include "opencv2/opencv.hpp"
