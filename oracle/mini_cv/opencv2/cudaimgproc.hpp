// mini_cv stand-in (TEST INFRASTRUCTURE ONLY): the cv::cuda:: slice used by the reference, served by the CPU functions of the same OpenCV
#include "mini_cv_cuda.hpp"
