{
  "targets": [{
    "target_name": "jsrt_addon",
    "sources": ["jsrt_addon.cc"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../jsraytracer_b200", "-ljsrt", "-Wl,-rpath,<(module_root_dir)/../../jsraytracer_b200"]
  }]
}
