// Minimal stand-in for <gnuradio/attributes.h>: symbol visibility macros.
#pragma once
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
