/* oracle/ref_harness/fm2d_f90_harness.c -- TEST INFRASTRUCTURE.  Driver around the mechanical translation of the reference's
 * fm2d/fm2d_ttime.f90 (oracle/f90toc.py -> oracle/_ref/fm2d_ttime_f2c.c, included below so that the module variables, which
 * the translation keeps file-local, can be set here).  It plays the part of modrays' ALLOCATE statements and assignments to
 * the variables of module globalp (fm2dray_cartesian.f90:180-260), then calls `travel` once.  Arrays are the caller's, column
 * major with leading dimension nnz, exactly as the Fortran holds veln / ttn / nsts (nnz, nnx). */
#include FM2D_F2C_SOURCE

int ref_fm2d_travel(int nnx_, int nnz_, double gox_, double goz_, double dnx_, double dnz_, int fom_, const double* veln_, double* ttn_,
                    int* nsts_, int urg, int vnl_, int vnr_, int vnt_, int vnb_, double scx, double scz, int* heap_pxpz, int* ntr_out) {
  nnx = nnx_; nnz = nnz_; gox = gox_; goz = goz_; dnx = dnx_; dnz = dnz_; fom = fom_;
  vnl = vnl_; vnr = vnr_; vnt = vnt_; vnb = vnb_;
  earth = 6371.0; /* (only feeds two dead assignments of fouds1 / fouds2) */
  veln = (double*)veln_; veln_d1 = nnz_; veln_d2 = nnx_; veln_l1 = 1; veln_l2 = 1;
  ttn = ttn_; ttn_d1 = nnz_; ttn_d2 = nnx_;
  nsts = nsts_; nsts_d1 = nnz_; nsts_d2 = nnx_;
  btg_d1 = nnx_ * nnz_ + 1; /* the Fortran allocates maxbt = NINT(snb*nnx*nnz) entries and never checks; every node fits here */
  btg = (backpointer*)calloc((size_t)btg_d1, sizeof(backpointer));
  ntr = 0;
  f90_stopped = 0;
  travel_(&scx, &scz, &urg);
  for (int i = 0; i < ntr; ++i) { heap_pxpz[2 * i] = btg[i].px; heap_pxpz[2 * i + 1] = btg[i].pz; }
  *ntr_out = ntr;
  free(btg);
  return f90_stopped;
}

/* bilinear alone: nv(2,2) column-major */
double ref_fm2d_bilinear(double dnx_, double dnz_, const double* nv, double dsx, double dsz) {
  double biv = 0.0;
  dnx = dnx_; dnz = dnz_;
  bilinear_((void*)nv, &dsx, &dsz, &biv);
  return biv;
}

/* gridder: velvin (nvz+2, nvx+2) column major -> the propagation grid veln (nnz, nnx), copied to veln_out.  The module's velv
 * (0:nvz+1, 0:nvx+1) stays allocated for ref_fm2d_bsplrefine. */
int ref_fm2d_gridder(int nvx_, int nvz_, int gdx_, int gdz_, const double* velvin, double* veln_out) {
  double go = 0.0, dv = 1.0;
  gdx = gdx_; gdz = gdz_;
  veln = 0;
  gridder_(&nvx_, &nvz_, &go, &go, &dv, &dv, (void*)velvin);
  for (int i = 0; i < nnx * nnz; ++i) veln_out[i] = veln[i];
  free(veln); veln = 0;
  return 0;
}

/* bsplrefine on the window (vnl..vnr, vnt..vnb) of the propagation grid; the refined extents are modrays' (nnx, nnz at the
 * time of the call: fm2dray_cartesian.f90:296-312), veln_out (nnzr, nnxr) starts as zeros like a fresh ALLOCATE here */
int ref_fm2d_bsplrefine(int nvx_, int nvz_, int gdx_, int gdz_, int sgdl_, int vnl_, int vnr_, int vnt_, int vnb_, const double* velvin,
                        int nnxr_, int nnzr_, double* veln_out) {
  double go = 0.0, dv = 1.0;
  gdx = gdx_; gdz = gdz_;
  veln = 0;
  gridder_(&nvx_, &nvz_, &go, &go, &dv, &dv, (void*)velvin); /* sets nvx, nvz and the module's velv as modrays has them */
  free(veln);
  sgdl = sgdl_; vnl = vnl_; vnr = vnr_; vnt = vnt_; vnb = vnb_;
  nnx = nnxr_; nnz = nnzr_;
  veln = veln_out; veln_d1 = nnzr_; veln_d2 = nnxr_; veln_l1 = 1; veln_l2 = 1;
  bsplrefine_();
  veln = 0;
  return 0;
}

/* srtimes for one source (csid = 1 of nsrc = 1): srs_ (nrc), ttime (nrc) */
int ref_fm2d_srtimes(int nnx_, int nnz_, double gox_, double goz_, double dnx_, double dnz_, const double* veln_, const double* ttn_,
                     double scx, double scz, int nrc_, const double* rcx_, const double* rcz_, const int* srs_, double* ttime) {
  int csid = 1, nsrc = 1;
  nnx = nnx_; nnz = nnz_; gox = gox_; goz = goz_; dnx = dnx_; dnz = dnz_;
  veln = (double*)veln_; veln_d1 = nnz_; veln_d2 = nnx_; veln_l1 = 1; veln_l2 = 1;
  ttn = (double*)ttn_; ttn_d1 = nnz_; ttn_d2 = nnx_;
  nrc = nrc_;
  rcx = (double*)rcx_; rcx_d1 = nrc_; rcz = (double*)rcz_; rcz_d1 = nrc_;
  srs = (int*)srs_; srs_d1 = nrc_; srs_d2 = 1;
  f90_stopped = 0;
  srtimes_(&scx, &scz, &csid, ttime, &nsrc);
  veln = 0;
  return f90_stopped;
}
