"""wlsqm_b200.utils -- mirrors the module layout of the reference's ``wlsqm.utils`` package."""
