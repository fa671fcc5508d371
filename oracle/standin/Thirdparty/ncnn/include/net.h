// oracle/standin/Thirdparty/ncnn/include/net.h (TEST INFRASTRUCTURE): the perfect tree's include/Detector.h -- pulled in by
// its include/KeyFrame.h -- names two ncnn types by pointer only; ncnn itself is absent and out of scope (the detector).
#ifndef B200_STANDIN_NCNN_NET_H
#define B200_STANDIN_NCNN_NET_H
namespace ncnn {
class Net;
class Mat;
}  // namespace ncnn
#endif
