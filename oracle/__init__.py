"""CPU oracle: test infrastructure only (see oracle/qso.h). Never imported by the product."""
