#ifndef CMCB200_SHIM_GSL_ERRNO_H
#define CMCB200_SHIM_GSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1, GSL_ERANGE = 2, GSL_EFAULT = 3,
       GSL_EINVAL = 4, GSL_EFAILED = 5, GSL_EFACTOR = 6, GSL_ESANITY = 7, GSL_ENOMEM = 8, GSL_EBADFUNC = 9,
       GSL_ERUNAWAY = 10, GSL_EMAXITER = 11, GSL_EZERODIV = 12, GSL_EBADTOL = 13, GSL_ETOL = 14,
       GSL_EUNDRFLW = 15, GSL_EOVRFLW = 16, GSL_ELOSS = 17, GSL_EROUND = 18, GSL_ESING = 21 };
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h);
gsl_error_handler_t *gsl_set_error_handler_off(void);
const char *gsl_strerror(const int gsl_errno);
#endif
