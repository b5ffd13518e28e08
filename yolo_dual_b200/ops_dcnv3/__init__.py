"""Drop-in for the reference's ``models/ops_dcnv3`` package (functions/, modules/)."""
