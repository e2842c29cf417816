// ORACLE shim: Viewer / MapDrawer headers sit on the include chain of Tracking.h; only forward declarations are needed to parse them.
#ifndef ORACLE_PANGOLIN_STUB
#define ORACLE_PANGOLIN_STUB
namespace pangolin { struct OpenGlMatrix { double m[16]; }; struct OpenGlRenderState; struct View; }
#endif
