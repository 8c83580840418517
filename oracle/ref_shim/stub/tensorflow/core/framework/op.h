// stand-in: the reference CPU functors include this header but use nothing from it
#include "tensorflow/core/framework/op_kernel.h"
