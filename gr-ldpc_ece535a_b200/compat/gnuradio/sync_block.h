// Minimal stand-in for <gnuradio/sync_block.h>: gr::sync_block lives in block.h here.
#pragma once
#include <gnuradio/block.h>
