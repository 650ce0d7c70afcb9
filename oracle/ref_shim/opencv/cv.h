#pragma once
// stand-in: the reference includes <opencv/cv.h> and uses nothing from it
