# Drop-in replacement for the worker body of splash.grid() (reference R/splash.grid.R:277-308).
# NOT RUN HERE (no R in this image).  Everything before and after clFun in splash.grid() stays as it is:
# terrain derivation (:77-110), the time axis (:145-167), the output bricks (:173-255) and the
# writeValues loop (:359-400) already work on [cells x layers] matrices per block of rows.

clFun <- function(i) {
	nr <- bs$nrows[i]
	# getValues() already returns [cells x days]: day-major with cells contiguous = the ABI layout
	sw  <- raster::getValues(sw_in, bs$row[i], nr)
	tcm <- raster::getValues(tc,    bs$row[i], nr)
	pnm <- raster::getValues(pn,    bs$row[i], nr)
	tt  <- as.Date(ztime)
	res <- .Call("splash_grid_run_R",
		sw, tcm, pnm,
		as.numeric(raster::getValues(lat,  bs$row[i], nr)),
		as.numeric(raster::getValues(elev, bs$row[i], nr)),
		as.numeric(raster::getValues(terraines[[1]], bs$row[i], nr)),
		as.numeric(raster::getValues(terraines[[2]], bs$row[i], nr)),
		raster::getValues(soil, bs$row[i], nr),            # [cells x 6]
		as.matrix(raster::getValues(Au, bs$row[i], nr)),   # [cells x 1] or [cells x 3]
		as.numeric(raster::getValues(resolution, bs$row[i], nr)),
		as.integer(format(tt, "%Y")), as.integer(format(tt, "%j")), as.integer(format(tt, "%m")),
		isTRUE(sim.control$monthly_out), 0L,
		PACKAGE = "rsplash")
	# the master loop indexes value[k, ] per cell and rbind()s them (:370-378); hand it the matrices
	# directly instead: writeValues(sm, res$wn, bs$row[b]) etc.
	res
}
