"""``isce3.ext`` namespace mirror so tests can read like the reference's
(``import isce3.ext.isce3 as isce`` -> ``import isce3_b200.ext.isce3 as isce``)."""
