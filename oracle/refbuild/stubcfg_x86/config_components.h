/* stand-in for the generated config_components.h: only the VVC decoder is enabled (see config.h) */
