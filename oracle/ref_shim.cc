// ref_shim.cc -- TEST INFRASTRUCTURE.  Reaches the delay-estimator knobs of an UNMODIFIED reference instance.
//
// The reference wires its delay estimator with robust validation off and max_lookahead 0 (aecm_core.cc:218,225) and
// offers no public call to change that; the estimator's own setters (delay_estimator_wrapper.h) need the handle that
// sits inside the instance.  AecMobile is a type private to echo_control_mobile.cc, so this translation unit INCLUDES
// that file from where it lies in the reference tree (nothing is copied; oracle/Makefile builds
// oracle/_ref/libaecm_ref_shim.so from this file plus the other, untouched reference objects, and only where
// /root/reference exists) and adds a handful of exported helpers that call the reference's own functions on the
// instance's estimator.  Used by tests (SURVEY 8-f4 parity) through tests/oracle_lib.py; never by the product.
#include "aecm/echo_control_mobile.cc"  // -I$(REF): the reference's wrapper, verbatim, with its file-local AecMobile
#include "aecm/delay_estimator_wrapper.h"

extern "C" {

static AecmCore *core_of(void *inst) { return inst ? static_cast<AecMobile *>(inst)->aecmCore : NULL; }

// WebRtc_enable_robust_validation / WebRtc_set_allowed_offset / WebRtc_set_lookahead on the instance's estimator.
// max_lookahead >= 0 first replaces the estimator by WebRtc_CreateDelayEstimator(farend, max_lookahead) and
// initialises it (what aecm_core.cc:217-225 + :391 do with 0).
int ref_shim_set_delay_estimator_options(void *inst, int robust, int allowed_offset, int max_lookahead, int lookahead) {
    AecmCore *c = core_of(inst);
    if (!c) return -1;
    if (max_lookahead >= 0) {
        void *fresh = WebRtc_CreateDelayEstimator(c->delay_estimator_farend, max_lookahead);
        if (!fresh) return -1;
        WebRtc_FreeDelayEstimator(c->delay_estimator);
        c->delay_estimator = fresh;
        if (WebRtc_InitDelayEstimator(c->delay_estimator) != 0) return -1;
    }
    if (lookahead >= 0 && WebRtc_set_lookahead(c->delay_estimator, lookahead) < 0) return -1;
    if (WebRtc_enable_robust_validation(c->delay_estimator, robust) != 0) return -1;
    return WebRtc_set_allowed_offset(c->delay_estimator, allowed_offset);
}

int ref_shim_soft_reset_delay_estimator(void *inst, int delay_shift) {
    AecmCore *c = core_of(inst);
    if (!c) return -1;
    WebRtc_SoftResetDelayEstimatorFarend(c->delay_estimator_farend, delay_shift);
    return WebRtc_SoftResetDelayEstimator(c->delay_estimator, delay_shift);
}

int ref_shim_last_delay(void *inst) {
    AecmCore *c = core_of(inst);
    return c ? WebRtc_last_delay(c->delay_estimator) : -1;
}

float ref_shim_last_delay_quality(void *inst) {
    AecmCore *c = core_of(inst);
    return c ? WebRtc_last_delay_quality(c->delay_estimator) : -1.f;
}

}  // extern "C"
