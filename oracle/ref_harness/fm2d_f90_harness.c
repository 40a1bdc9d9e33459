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

/* where rpaths' rays go (the translation calls this instead of filling the caller's T_RAY array: oracle/f90toc.py) */
static __thread int* g_ray_npts;
static __thread double* g_ray_pts;
static __thread int g_ray_cap, g_ray_slots, g_ray_err;
static void f90_ray_store(int slot, int nrp, const double* praypts, int csid, int revid) {
  (void)csid; (void)revid;
  if (!g_ray_npts) return;
  if (slot < 1 || slot > g_ray_slots || nrp > g_ray_cap) { g_ray_err = 1; return; }
  g_ray_npts[slot - 1] = nrp;
  for (int j = 0; j < nrp; ++j) {
    g_ray_pts[((size_t)(slot - 1) * g_ray_cap + j) * 2 + 0] = praypts[2 * j + 0];   /* praypts(1,j) */
    g_ray_pts[((size_t)(slot - 1) * g_ray_cap + j) * 2 + 1] = praypts[2 * j + 1];   /* praypts(2,j) */
  }
}

/* modrays (wrgf = 0, no timeField): the statements of fm2dray_cartesian.f90:137-205 and the skeleton of its source loop
 * (:209-212, :440-450) are written out here -- allocations and assignments of dummy arguments to module variables; everything
 * that computes is the translation: gridder, modrays_source (the loop body :213-438), srtimes and, with uar = 0 (group-velocity
 * data), rpaths.  velvin (nvz+2, nvx+2), srs_ / srsv_ (nrc, nsrc), ttime (nrc, nsrc), all column major; field: optional
 * (nnz, nnx, nsrc), the travel-time field of every marched source (ttn(1:nnz,1:nnx) as `timeField(:,:,i) = ttn(1:nnz,1:nnx)`
 * would take it); rays: ray_npts[nrc*nsrc] (zeroed by the caller), ray_pts[nrc*nsrc][cap][2], by slot srsv - 1; *crazy =
 * modrays' crazyray.  Returns 0, 1 (STOP: a source outside the model), 3 (STOP: a receiver outside the model), 4 (a ray that
 * does not fit the caller's slots). */
int ref_fm2d_modrays(int nsrc_, const double* scx_, const double* scz_, int nrc_, const double* rcx_, const double* rcz_,
                     const int* srs_, int nvx_, int nvz_, double gox_, double goz_, double dvx_, double dvz_, const double* velvin,
                     int gdx_, int gdz_, int asgr_, int sgdl_, int sgs_, int fom_, double snb_, double* ttime, double* field,
                     int uar, const int* srsv_, int cap, int* ray_npts, double* ray_pts, int* crazy) {
  int rc = 0, crazyray = 0;
  nrc = nrc_;
  rcx = (double*)rcx_; rcx_d1 = nrc_; rcz = (double*)rcz_; rcz_d1 = nrc_;
  srs = (int*)srs_; srs_d1 = nrc_; srs_d2 = nsrc_;
  srsv = (int*)srsv_; srsv_d1 = nrc_; srsv_d2 = nsrc_;
  gdx = gdx_; gdz = gdz_; asgr = asgr_; sgdl = sgdl_; fom = fom_; earth = 6371.0; snb = snb_;
  veln = 0; velnb = 0; ttnr = 0; nstsr = 0; npts = 0; raypts = 0;
  g_ray_npts = ray_npts; g_ray_pts = ray_pts; g_ray_cap = cap; g_ray_slots = nrc_ * nsrc_; g_ray_err = 0;
  gridder_(&nvx_, &nvz_, &gox_, &goz_, &dvx_, &dvz_, (void*)velvin);
  const int cnx = nnx, cnz = nnz;
  ttn_l1 = ttn_l2 = 1; ttn_d1 = nnz; ttn_d2 = nnx; ttn = (double*)calloc((size_t)nnz * nnx, sizeof(double));   /* ALLOCATE(ttn(nnz,nnx)) */
  nsts_l1 = nsts_l2 = 1; nsts_d1 = nnz; nsts_d2 = nnx; nsts = (int*)calloc((size_t)nnz * nnx, sizeof(int));   /* ALLOCATE(nsts(nnz,nnx)) */
  btg_l1 = 1; btg_d1 = (int)lround(snb * nnx * nnz); btg = (backpointer*)calloc((size_t)btg_d1 + 64, sizeof(backpointer)); /* maxbt=NINT(snb*nnx*nnz); ALLOCATE(btg(maxbt)) (+ the translator's pad) */
  f90_stopped = 0;
  for (int i = 1; i <= nsrc_ && !rc; ++i) {
    int any = 0, nsrc = nsrc_, sgs = sgs_, wrgf = 0;
    double x = 0, z = 0;
    for (int r = 0; r < nrc_; ++r) any += srs_[(size_t)(i - 1) * nrc_ + r];
    if (any == 0 && i != 1) continue;                     /* IF(SUM(srs(:,i)).EQ.0.AND.i.NE.1) cycle */
    modrays_source_(&i, &nsrc, &sgs, (void*)scx_, (void*)scz_, &x, &z);
    if (f90_stopped) { rc = 1; break; }
    if (field)
      for (int k = 0; k < cnx; ++k)
        for (int j = 0; j < cnz; ++j) field[((size_t)(i - 1) * cnx + k) * cnz + j] = ttn[(size_t)k * ttn_d1 + j];
    srtimes_(&x, &z, &i, ttime, &nsrc);
    if (f90_stopped) { rc = 3; break; }
    if (uar == 0) {                                       /* IF(uar .eq. 0 .OR. wrgf.eq.i.OR.wrgf.LT.0)THEN */
      rpaths_(&wrgf, &uar, &i, &x, &z, 0);
      if (f90_stopped) { rc = 3; break; }
      if (g_ray_err) { rc = 4; break; }
      crazyray = crazyray + crazyrp;
    }
    if (asgr == 1 && ttnr) { free(ttnr); ttnr = 0; free(nstsr); nstsr = 0; }   /* IF(asgr.EQ.1 .and. allocated(ttnr)) DEALLOCATE(ttnr,nstsr) */
  }
  if (crazy) *crazy = crazyray;
  free(velnb); velnb = 0; free(ttnr); ttnr = 0; free(nstsr); nstsr = 0;
  free(veln); veln = 0; free(ttn); ttn = 0; free(nsts); nsts = 0; free(btg); btg = 0; free(velv); velv = 0;
  g_ray_npts = 0;
  return rc;
}
int ref_fm2d_modrays_times(int nsrc_, const double* scx_, const double* scz_, int nrc_, const double* rcx_, const double* rcz_,
                           const int* srs_, int nvx_, int nvz_, double gox_, double goz_, double dvx_, double dvz_, const double* velvin,
                           int gdx_, int gdz_, int asgr_, int sgdl_, int sgs_, int fom_, double snb_, double* ttime, double* field) {
  return ref_fm2d_modrays(nsrc_, scx_, scz_, nrc_, rcx_, rcz_, srs_, nvx_, nvz_, gox_, goz_, dvx_, dvz_, velvin, gdx_, gdz_, asgr_, sgdl_,
                          sgs_, fom_, snb_, ttime, field, 1, srs_, 0, 0, 0, 0);
}
