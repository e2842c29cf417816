// ORACLE shim: shadows the reference's g2o header (which needs Eigen) for translation units that only see Converter.h declarations.
#ifndef ORACLE_G2O_STUB
#define ORACLE_G2O_STUB
namespace g2o { class SE3Quat; struct Sim3; }
#endif
