"""CPU: the numpy restatement of the terrain slice (oracle/terrain_oracle.py; parity unpinned, see its header) behaves as
the in-tree R code says: ncellflow's template points every neighbour at the centre, zero matches count as one, an all-NA
window is NA; slopes of valid cells that terrain() leaves NA (borders, NA neighbours) are 0; a tilted plane has the
analytic slope and aspect."""
import numpy as np

from oracle import terrain_oracle as to


def test_ncellflow_counts_the_neighbours_that_drain_into_the_centre():
    fd = np.full((5, 5), np.nan)
    fd[1:4, 1:4] = [[2, 4, 8], [1, 64, 16], [128, 64, 32]]   # all eight neighbours of (2, 2) point at it
    nin = to.ncellflow(fd, "in")
    assert nin[2, 2] == 8
    assert np.isnan(to.ncellflow(np.full((4, 4), np.nan), "in")).all()
    lone = np.full((3, 3), np.nan)
    lone[1, 1] = 4.0
    assert to.ncellflow(lone, "in")[1, 1] == 1 and to.ncellflow(lone, "out")[1, 1] == 1   # nmatch[nmatch == 0] <- 1
    # 'out' reverses the template: the centre's own code never matters (template 0), neighbours pointing AWAY count
    away = np.full((3, 3), np.nan)
    away[:, :] = [[32, 64, 128], [16, 0, 1], [8, 4, 2]]
    # (the centre is compared as well: the template's centre is 0, and so is this window's -> one more match, as in R)
    assert to.ncellflow(away, "out")[1, 1] == 9 and to.ncellflow(away, "in")[1, 1] == 1


def test_tilted_plane_has_the_analytic_slope_aspect_and_flow_direction():
    nr, nc = 12, 15
    x = np.arange(nc)[None, :] * 100.0
    y = -np.arange(nr)[:, None] * 100.0          # north row first: y decreases downwards
    z = 500.0 + 0.1 * x + 0.0 * y                # rises to the east: faces west (270), drains west (code 16)
    t = to.terrain(z, ymax=0.0, xres=100.0, yres=100.0, lonlat=False)
    inner = (slice(1, -1), slice(1, -1))
    assert np.allclose(t["slope"][inner], np.degrees(np.arctan(0.1)), rtol=1e-12)
    assert np.allclose(t["aspect"][inner], 270.0) and (t["flowdir"][inner] == 16).all()
    assert (t["slope"][0] == 0).all() and np.isnan(t["flowdir"][0]).all()          # border: NA -> 0 for slope, NA flowdir
    z2 = 500.0 + 0.05 * y                         # rises to the north: faces south (180), drains south (code 4)
    t2 = to.terrain(z2 + 0 * x, ymax=0.0, xres=100.0, yres=100.0, lonlat=False)
    assert np.allclose(t2["aspect"][inner], 180.0) and (t2["flowdir"][inner] == 4).all()
    z[5, 7] = np.nan
    t3 = to.terrain(z, ymax=0.0, xres=100.0, yres=100.0, lonlat=False)
    assert np.isnan(t3["slope"][5, 7]) and t3["slope"][5, 8] == 0.0 and np.isnan(t3["lat"][5, 7]) and not np.isnan(t3["lat"][5, 8])


def test_geographic_cell_sizes_shrink_with_latitude():
    lat, dx, dy = to.cell_sizes(4, ymax=60.0, xres=0.5, yres=0.5)
    assert np.allclose(lat, [59.75, 59.25, 58.75, 58.25]) and np.allclose(dy, 6378137.0 * np.radians(0.5))
    assert np.all(np.diff(dx) > 0) and np.isclose(dx[0], 6378137.0 * np.cos(np.radians(59.75)) * np.radians(0.5))
