# Drop-in replacement for the worker body of splash.grid() (reference R/splash.grid.R:277-308).
# NOT RUN HERE (no R in this image).  Everything before and after clFun in splash.grid() stays as it is:
# terrain derivation (:77-110), the time axis (:145-167), the output bricks (:173-255) and the
# writeValues loop (:359-400) already work on [cells x layers] matrices per block of rows.
#
# Before library(rsplash):  Sys.setenv(CUDA_DEVICE_MAX_CONNECTIONS = "32")   (see INTEGRATION.md)
# `devices`: integer vector of CUDA ordinals, e.g. 0L or 0:7 (a block is then spread over all of them).

dbl <- function(x) { storage.mode(x) <- "double"; x }   # integer rasters (INT2S DEMs, ...) must not reach REAL()

block_args <- function(i) {
	nr <- bs$nrows[i]
	tt <- as.Date(ztime)
	# getValues() already returns [cells x days]: day-major with cells contiguous = the ABI layout
	list(dbl(raster::getValues(sw_in, bs$row[i], nr)),
		dbl(raster::getValues(tc,    bs$row[i], nr)),
		dbl(raster::getValues(pn,    bs$row[i], nr)),
		as.numeric(raster::getValues(lat,  bs$row[i], nr)),
		as.numeric(raster::getValues(elev, bs$row[i], nr)),
		as.numeric(raster::getValues(terraines[[1]], bs$row[i], nr)),
		as.numeric(raster::getValues(terraines[[2]], bs$row[i], nr)),
		dbl(raster::getValues(soil, bs$row[i], nr)),            # [cells x 6]
		dbl(as.matrix(raster::getValues(Au, bs$row[i], nr))),   # [cells x 1] or [cells x 3]
		as.numeric(raster::getValues(resolution, bs$row[i], nr)),
		as.integer(format(tt, "%Y")), as.integer(format(tt, "%j")), as.integer(format(tt, "%m")),
		isTRUE(sim.control$monthly_out))
}

# (a) synchronous: one call per block, as clFun was
clFun <- function(i, devices = 0L) {
	res <- do.call(.Call, c(list("splash_grid_run_R"), block_args(i), list(as.integer(devices)), list(PACKAGE = "rsplash")))
	# the master loop indexes value[k, ] per cell and rbind()s them (:370-378); hand it the matrices
	# directly instead: writeValues(sm, res$wn, bs$row[b]) etc.
	res
}

# (b) the reference's own scheduling loop (:312-314, 359-400) with sendCall / recvOneData replaced:
#     keep `lanes` blocks in flight per GPU, write each block away as it comes back
run_blocks <- function(devices = 0L, lanes = 2L, write_block) {
	block_of <- list()
	submit <- function(i) {
		tk <- do.call(.Call, c(list("splash_grid_submit_R"), block_args(i), list(as.integer(devices), as.integer(lanes)), list(PACKAGE = "rsplash")))
		block_of[[as.character(tk)]] <<- i
	}
	nxt <- 1L
	while (nxt <= min(bs$n, length(devices) * lanes)) { submit(nxt); nxt <- nxt + 1L }   # sendCall to every worker, :312-314
	for (done in seq_len(bs$n)) {
		d <- .Call("splash_grid_wait_R", -1, PACKAGE = "rsplash")                          # recvOneData, :361
		write_block(block_of[[as.character(d$ticket)]], d$value)                          # writeValues / matrix fill, :370-391
		if (nxt <= bs$n) { submit(nxt); nxt <- nxt + 1L }                                   # sendCall of the next block, :397
	}
}
