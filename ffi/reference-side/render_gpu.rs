//! Reference-side patch, part 3: the one new method next to `Camera::render` (src/camera.rs:79-126).  Every scene function
//! of src/main.rs then calls `camera.render_gpu(&world, "demo/….png")` instead of `camera.render(…)`; nothing else changes.
use image::{ImageBuffer, Rgb};
use pt_b200_sys::*;

use crate::{camera::Camera, flatten, hittable::World};

fn check(rc: std::os::raw::c_int) { if rc != PT_OK { panic!("pt_b200 error {rc}: {}", last_error()); } }

impl Camera {
    /// `gpus` = 1: pt_render on device 0; more: pt_render_multi (samples split over the devices, reduce on the device).
    pub fn render_gpu(&self, world: &World, filename: &str, gpus: usize) {
        let flat = flatten::flatten(world, &self.environment);
        let cam = pt_camera {
            aspect_ratio: self.aspect_ratio, image_width: self.image_width as u32, samples_per_pixel: self.samples_per_pixel as u32,
            max_depth: self.max_depth as u32, env_is_map: flat.env_image.is_some() as u32, vfov: self.vfov,
            look_from: pt_vec3 { x: self.look_from.x, y: self.look_from.y, z: self.look_from.z },
            look_at: pt_vec3 { x: self.look_at.x, y: self.look_at.y, z: self.look_at.z },
            vup: pt_vec3 { x: self.vup.x, y: self.vup.y, z: self.vup.z },
            blur_strength: self.blur_strength, focal_length: self.focal_length, defocus_angle: self.defocus_angle,
            env_color: flat.env_color, env_image: flat.env_image.unwrap_or(PT_NONE), _pad: 0,
        };
        let params = pt_render_params { seed: rand::random(), sample_begin: 0, sample_count: cam.samples_per_pixel, sample_stride: 1,
                                        nan_policy: PT_NAN_REFERENCE, pool_paths: 0, flags: 0 };
        let desc = flat.desc();
        let h = unsafe { pt_camera_image_height(&cam) } as usize;
        let mut mean = vec![0f32; self.image_width * h * 3];
        let mut stats = pt_stats::default();
        println!("rendering production");                                                     // camera.rs:101
        unsafe {
            if gpus > 1 {
                let devices: Vec<i32> = (0..gpus as i32).collect();
                check(pt_render_multi(gpus as i32, devices.as_ptr(), &desc, &cam, &params, mean.as_mut_ptr(), &mut stats));
            } else {
                let (mut ctx, mut scene) = (std::ptr::null_mut(), std::ptr::null_mut());
                check(pt_ctx_create(0, &mut ctx));
                check(pt_scene_create(ctx, &desc, &mut scene));
                check(pt_render(ctx, scene, &cam, &params, mean.as_mut_ptr(), &mut stats));
                pt_scene_destroy(scene);
                pt_ctx_destroy(ctx);
            }
        }
        let mut img = ImageBuffer::<Rgb<u8>, Vec<u8>>::new(self.image_width as u32, h as u32);
        for (i, px) in img.pixels_mut().enumerate() {                                         // camera.rs:109-114,128-130
            let g = |x: f32| ((x as f64).max(0.0).sqrt().clamp(0.0, 0.999) * 256.0) as u8;
            *px = Rgb([g(mean[3 * i]), g(mean[3 * i + 1]), g(mean[3 * i + 2])]);
        }
        if let Err(err) = img.save(filename) { eprintln!("Failed to save image {err}"); }     // camera.rs:118-123
        eprintln!("[pt_b200] {} paths, {} segments, {:.3} s on the device", stats.paths, stats.segments, stats.device_ms * 1e-3);
    }
}
