/* -*- c++ -*- */
/*
 * Export macro of the ldpc_ece535a module -- same name and meaning as the reference's
 * include/ldpc_ece535a/api.h:27-31, so sources written against it compile unchanged.
 */
#ifndef INCLUDED_LDPC_ECE535A_API_H
#define INCLUDED_LDPC_ECE535A_API_H

#include <gnuradio/attributes.h>

#ifdef gnuradio_ldpc_ece535a_EXPORTS
#define LDPC_ECE535A_API __GR_ATTR_EXPORT
#else
#define LDPC_ECE535A_API __GR_ATTR_IMPORT
#endif

#endif /* INCLUDED_LDPC_ECE535A_API_H */
