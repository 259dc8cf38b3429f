"""Drop-in layout for RichSem's ``models/richsem/ops`` package (functions/, modules/)."""
