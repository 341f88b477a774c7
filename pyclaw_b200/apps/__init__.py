"""Application-level helpers that the reference ships as per-app Fortran (apps/*/)."""
