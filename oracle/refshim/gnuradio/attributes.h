// Stand-in for <gnuradio/attributes.h> (oracle/_ref build only): symbol visibility macros.
#pragma once
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
