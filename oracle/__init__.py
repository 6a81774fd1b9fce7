"""oracle/ -- TEST INFRASTRUCTURE.  CPU restatement of the YALPS hot path used only as the
parity checker (tests/, __graft_entry__.smoke(), bench.py CPU legs).  Never imported by yalps_b200."""
