#pragma once
#include <ros/ros.h>
namespace sensor_msgs { struct Imu { std_msgs::Header header; typedef std::shared_ptr<Imu const> ConstPtr; }; }
