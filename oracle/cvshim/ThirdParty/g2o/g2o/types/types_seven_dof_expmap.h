// ORACLE shim: shadows the reference's g2o header (which needs Eigen) for translation units that only see declarations over these
// types (Converter.h, Optimizer.h) or hold them as never-touched members (LoopClosing.h). Opaque placeholders, no arithmetic.
#ifndef ORACLE_G2O_STUB
#define ORACLE_G2O_STUB
namespace g2o { class SE3Quat { double opaque_[8]; }; struct Sim3 { double opaque_[8]; }; }
#endif
