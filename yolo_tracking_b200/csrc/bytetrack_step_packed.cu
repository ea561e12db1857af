// The ByteTrack / BoT-SORT frame step instantiated for the packed frame interface (b200track_step_packed): detection rows
// of all streams back to back (fp32 or fp64), compact result rows at the same offsets.  Same kernel source as
// bytetrack_step.cu; see the note at its top.
#define B200_STEP_PACKED 1
#include "bytetrack_step.cu"
