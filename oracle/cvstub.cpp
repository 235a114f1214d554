// TEST INFRASTRUCTURE ONLY (oracle/): satisfies the 11 undefined cv::* symbols of the reference's
// shipped binary Segmentation/Segmentation/cython/src/liblpbox_solver.so so that it can be
// dlopen()ed without OpenCV.  None of these are reached by the entry points the harness drives
// (ADMM_bqp_linear_ineq / ADMM_bqp_unconstrained / _conjugate_gradient / mat_mul_vec / graph
// builder helpers); every stub aborts loudly if it is ever called.
#include <cstdio>
#include <cstdlib>
#define STUB(sym) extern "C" void sym() { std::fprintf(stderr, "cvstub: %s called\n", #sym); std::abort(); }
STUB(_ZN2cv3Mat10deallocateEv)
STUB(_ZN2cv3Mat20updateContinuityFlagEv)
STUB(_ZN2cv3Mat6createEiPKii)
STUB(_ZN2cv5errorEiRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEPKcS9_i)
STUB(_ZN2cv6imreadERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEi)
STUB(_ZN2cv6resizeERKNS_11_InputArrayERKNS_12_OutputArrayENS_5Size_IiEEddi)
STUB(_ZN2cv7imwriteERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERKNS_11_InputArrayERKSt6vectorIiSaIiEE)
STUB(_ZN2cv8fastFreeEPv)
STUB(_ZN2cv9transposeERKNS_11_InputArrayERKNS_12_OutputArrayE)
STUB(_ZNK2cv3Mat1tEv)
STUB(_ZNK2cv3Mat9convertToERKNS_12_OutputArrayEidd)
