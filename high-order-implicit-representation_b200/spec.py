"""Constants of the reference dependency ``high-order-layers-torch ^2.4.2`` that SURVEY.md
Appendix A records as recalled rather than verified.  They live here (product side) and at the
top of ``oracle/hol_oracle.py`` (checker side) only, so that a correction after A.9 is a
one-line change on each side; every kernel takes them as descriptor fields."""

#: Fourier basis argument is FOURIER_ARG_SCALE * pi * i * x / length  (A.5)
FOURIER_ARG_SCALE = 2.0
#: odd j -> sin, even j -> cos  (A.5)
FOURIER_ODD_IS_SIN = True
#: MaxAbsNormalization eps  (A.2)
MAXABS_EPS = 1e-6
#: positions_from_mesh normalisation (A.6): "max_size" -> (v/max(width,height) - 0.5)*2,
#: "size_minus_1" -> v/(size-1)*2-1
MESH_NORMALIZE = "max_size"
