// Keras HDF5 weight file front end (models/vgg2_mobilenet.h5; facerec_test.py:326-334).  Placeholder until the
// hand-written HDF5 reader lands: fails loudly instead of guessing.
#include "graph.h"

namespace hfr {

Plan compile_keras_mobilenet_h5(const uint8_t*, size_t, int) {
  throw std::runtime_error("hdf5: Keras .h5 loading is not implemented in this build; convert to a frozen .pb");
}

}  // namespace hfr
