"""wlsqm_b200.fitter -- mirrors the module layout of the reference's ``wlsqm.fitter`` package."""
