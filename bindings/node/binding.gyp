{
  "targets": [{
    "target_name": "yalps_b200",
    "sources": ["addon.c"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../yalps_b200", "-lyalps_b200", "-Wl,-rpath,<(module_root_dir)/../../yalps_b200"]
  }]
}
