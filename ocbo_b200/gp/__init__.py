"""GP model interface of the reference (get_gp_fitter -> fitter -> GP) over the B200 engine."""
