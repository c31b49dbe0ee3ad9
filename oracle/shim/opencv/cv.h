// shim (test infrastructure): scanRegistration.cpp:43 includes <opencv/cv.h> and uses nothing from it
