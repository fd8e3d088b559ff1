"""ORACLE: CPU restatement of the reference algorithm. Test infrastructure only --
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by
the product path (fv3-jedi-linearmodel_b200/)."""
