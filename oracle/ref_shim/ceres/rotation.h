#pragma once
// stand-in: the reference includes <ceres/rotation.h> and uses nothing from it
