{
  "targets": [{
    "target_name": "msm_b200",
    "sources": ["msm_b200_addon.c"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../msm_zprize_b200/csrc", "-lmsm_b200",
                  "-Wl,-rpath,<(module_root_dir)/../msm_zprize_b200/csrc"],
    "cflags": ["-O2", "-Wall"]
  }]
}
