// Test infrastructure (oracle/ref_shim_tracking): the two OpenCV types adapter/tracking_gpba.h reads through the reference's
// MultiFrame (cv::KeyPoint::pt / octave).
#pragma once
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
}  // namespace cv
